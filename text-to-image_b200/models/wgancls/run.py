"""Mirror of the reference's ``models/wgancls/run.py`` (:19-70): ``--cfg`` YAML, output directories, dataset, then the
mode dispatch on ``cfg.EVAL.FLAG`` / ``cfg.TRAIN.FLAG``.  Training runs ``WGanClsTrainer.train()`` on the B200 model.
The evaluation (Inception score / FID, ``eval_wgan.py``) and visualisation (``visualize_wgan.py``) consumers are out of
scope for this path: their modes raise ``NotImplementedError`` naming what is missing;
both only need ``WGanCls(cfg, build_model=False).generator(...)``, which this package provides.

The reference reads pickled datasets (``preprocess/dataset.py``: out of scope, no data in this repository): when
``cfg.DATASET_DIR`` has no ``train`` / ``test`` pickles the synthetic stand-in with the same two methods is used.

    python -m t2i_b200.models.wgancls.run --cfg text-to-image_b200/models/wgancls/cfg/flowers.yml [--max-steps N]
"""
import argparse
import os

from ...utils.config import config_from_yaml
from .model import WGanCls
from .trainer import SyntheticTextDataset, WGanClsTrainer


def load_dataset(cfg):
    """TextDataset(datadir, 64) + get_data(train / test) in the reference (:33-40)."""
    datadir = cfg.DATASET_DIR
    if os.path.exists('%s/train' % datadir) and os.path.exists('%s/test' % datadir):
        raise NotImplementedError("reading the reference's pickled datasets (preprocess/dataset.py) is out of scope; "
                                  "pass a dataset object with .train.next_batch / .test.next_batch_test")
    print(" [!] no dataset under %s: synthetic batches (images ~ U(-1, 1), embeddings ~ N(0, 1))" % datadir)
    return SyntheticTextDataset(embed_dim=cfg.MODEL.EMBED_DIM)


def main(cfg_path, dataset=None, max_steps=None, **model_kw):
    print(cfg_path)
    cfg = config_from_yaml(cfg_path)
    if max_steps is not None:
        cfg.TRAIN.MAX_STEPS = max_steps

    if not os.path.exists(cfg.CHECKPOINT_DIR):
        os.makedirs(cfg.CHECKPOINT_DIR)
    if not os.path.exists(cfg.SAMPLE_DIR):
        os.makedirs(cfg.SAMPLE_DIR)
    if not os.path.exists(cfg.LOGS_DIR):
        os.makedirs(cfg.LOGS_DIR)

    dataset = dataset if dataset is not None else load_dataset(cfg)

    if cfg.EVAL.FLAG:
        WGanCls(cfg, build_model=False, **model_kw)
        raise NotImplementedError("EVAL.FLAG: the Inception-score / FID evaluation (models/wgancls/eval_wgan.py, "
                                  "models/inception) is out of scope; sample with WGanCls(cfg, build_model=False).generator")
    elif cfg.TRAIN.FLAG:
        wgan = WGanCls(cfg, **model_kw)
        trainer = WGanClsTrainer(
            sess=None,
            model=wgan,
            dataset=dataset,
            cfg=cfg,
        )
        trainer.train()
        return trainer
    else:
        WGanCls(cfg, build_model=False, **model_kw)
        raise NotImplementedError("visualisation mode (models/wgancls/visualize_wgan.py) is out of scope")


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--cfg', default='./text-to-image_b200/models/wgancls/cfg/flowers.yml',
                    help='Relative path to the config of the model')
    ap.add_argument('--max-steps', type=int, default=None)
    a = ap.parse_args()
    main(a.cfg, max_steps=a.max_steps)
