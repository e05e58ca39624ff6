"""Mirror of the reference's ``models/stackgan/stageII/run.py`` (:26-86): the two YAML configs (stage I, stage II), the
output directories, the dataset, then the mode dispatch on ``cfg.EVAL.FLAG`` / ``cfg.TRAIN.FLAG``.  Training builds the
stage-I model with ``build_model=False`` (only its generator is used, :64), the stage-II model around it and runs
``ConditionalGanTrainer.train()``.  Evaluation / visualisation (``eval_stageii.py``, ``visualize_stageiI.py``) are out of
scope; the reference's pickled datasets likewise: without them the synthetic stand-in is used.

    python -m t2i_b200.models.stackgan.stageII.run --cfg_stage_I <yml> --cfg_stage_II <yml> [--max-updates N]
"""
import argparse
import os

from ....utils.config import config_from_yaml
from ...wgancls.trainer import SyntheticTextDataset
from ..stageI.model import ConditionalGan as ConditionalGanStageI
from .model import ConditionalGan
from .trainer import ConditionalGanTrainer


def main(cfg_stage_i_path, cfg_path, dataset=None, max_updates=None, **model_kw):
    cfg_stage_i = config_from_yaml(cfg_stage_i_path)
    cfg = config_from_yaml(cfg_path)

    if not os.path.exists(cfg.CHECKPOINT_DIR):
        os.makedirs(cfg.CHECKPOINT_DIR)
    if not os.path.exists(cfg.SAMPLE_DIR):
        os.makedirs(cfg.SAMPLE_DIR)
    if not os.path.exists(cfg.LOGS_DIR):
        os.makedirs(cfg.LOGS_DIR)

    if dataset is None:
        datadir = cfg.DATASET_DIR
        if os.path.exists('%s/train' % datadir) and os.path.exists('%s/test' % datadir):
            raise NotImplementedError("reading the reference's pickled datasets (preprocess/dataset.py) is out of scope; "
                                      "pass a dataset object with .train.next_batch / .test.next_batch_test")
        print(" [!] no dataset under %s: synthetic 256x256 batches" % datadir)
        dataset = SyntheticTextDataset(embed_dim=cfg.MODEL.EMBED_DIM, image_size=256)

    if cfg.EVAL.FLAG:
        stage_i = ConditionalGanStageI(cfg_stage_i, build_model=False, **model_kw)
        ConditionalGan(stage_i, cfg, build_model=False)
        raise NotImplementedError("EVAL.FLAG: the Inception-score evaluation (models/stackgan/stageII/eval_stageii.py) is "
                                  "out of scope; sample with ConditionalGan(stage_i, cfg).sample(z, embed)")
    elif cfg.TRAIN.FLAG:
        stage_i = ConditionalGanStageI(cfg_stage_i, build_model=False, **model_kw)
        stage_ii = ConditionalGan(stage_i, cfg)
        stage_ii_trainer = ConditionalGanTrainer(
            sess=None,
            model=stage_ii,
            dataset=dataset,
            cfg=cfg,
            cfg_stage_i=cfg_stage_i,
        )
        stage_ii_trainer.train(max_updates=max_updates)
        return stage_ii_trainer
    else:
        stage_i = ConditionalGanStageI(cfg_stage_i, build_model=False, **model_kw)
        ConditionalGan(stage_i, cfg, build_model=False)
        raise NotImplementedError("visualisation mode (models/stackgan/stageII/visualize_stageiI.py) is out of scope")


if __name__ == '__main__':
    here = os.path.dirname(os.path.abspath(__file__))
    ap = argparse.ArgumentParser()
    ap.add_argument('--cfg_stage_I', default=os.path.join(here, '..', 'stageI', 'cfg', 'flowers.yml'))
    ap.add_argument('--cfg_stage_II', default=os.path.join(here, 'cfg', 'flowers.yml'))
    ap.add_argument('--max-updates', type=int, default=None)
    a = ap.parse_args()
    main(a.cfg_stage_I, a.cfg_stage_II, max_updates=a.max_updates)
