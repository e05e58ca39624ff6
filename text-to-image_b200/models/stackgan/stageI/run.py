"""Mirror of the reference's ``models/stackgan/stageI/run.py``: ``--cfg`` YAML, output directories, dataset, then the mode
dispatch on ``cfg.EVAL.FLAG`` / ``cfg.TRAIN.FLAG``; training runs ``ConditionalGanTrainer.train()`` on the B200 model.
Evaluation / visualisation (``eval_stagei.py``, ``visualize_stagei.py``) and the pickled datasets are out of scope
(SURVEY.md section 2); without a dataset the synthetic stand-in is used.

    python -m t2i_b200.models.stackgan.stageI.run --cfg <yml> [--max-updates N]
"""
import argparse
import os

from ....utils.config import config_from_yaml
from ...wgancls.trainer import SyntheticTextDataset
from .model import ConditionalGan
from .trainer import ConditionalGanTrainer


def main(cfg_path, dataset=None, max_updates=None, **model_kw):
    cfg = config_from_yaml(cfg_path)
    for d in (cfg.CHECKPOINT_DIR, cfg.SAMPLE_DIR, cfg.LOGS_DIR):
        if not os.path.exists(d):
            os.makedirs(d)
    if dataset is None:
        datadir = cfg.DATASET_DIR
        if os.path.exists('%s/train' % datadir) and os.path.exists('%s/test' % datadir):
            raise NotImplementedError("reading the reference's pickled datasets (preprocess/dataset.py) is out of scope; "
                                      "pass a dataset object with .train.next_batch / .test.next_batch_test")
        print(" [!] no dataset under %s: synthetic 64x64 batches" % datadir)
        dataset = SyntheticTextDataset(embed_dim=cfg.MODEL.EMBED_DIM)
    if cfg.EVAL.FLAG:
        ConditionalGan(cfg, build_model=False, **model_kw)
        raise NotImplementedError("EVAL.FLAG: the Inception-score evaluation (models/stackgan/stageI/eval_stagei.py) is out of scope")
    elif cfg.TRAIN.FLAG:
        stage_i = ConditionalGan(cfg, **model_kw)
        stage_i_trainer = ConditionalGanTrainer(
            sess=None,
            model=stage_i,
            dataset=dataset,
            cfg=cfg,
        )
        stage_i_trainer.train(max_updates=max_updates)
        return stage_i_trainer
    else:
        ConditionalGan(cfg, build_model=False, **model_kw)
        raise NotImplementedError("visualisation mode (models/stackgan/stageI/visualize_stagei.py) is out of scope")


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--cfg', default=os.path.join(os.path.dirname(os.path.abspath(__file__)), 'cfg', 'flowers.yml'))
    ap.add_argument('--max-updates', type=int, default=None)
    a = ap.parse_args()
    main(a.cfg, max_updates=a.max_updates)
