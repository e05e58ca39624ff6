"""Mirror of the reference's ``models/pggan/train_pggan.py`` (:17-69): the 15-entry stage schedule (every stage > 1 is
run twice: a transition pass that fades the new layers in, then a stabilisation pass), batch 16 (8 from stage 6 on),
600 000 images per pass, checkpoints read from the previous entry's directory and written to this stage's."""
import os

from .pggan import PGGAN

STAGE = [1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8]          # train_pggan.py:17
PREV_STAGE = [1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8]     # :18


def schedule(images=600000):
    """(stage, previous stage, trans, batch size, iterations) of every pass (train_pggan.py:20-31)."""
    out = []
    for i in range(len(STAGE)):
        t = False if (i % 2 == 0) else True
        batch_size = 16
        if STAGE[i] >= 6:
            batch_size = 8
        out.append((STAGE[i], PREV_STAGE[i], t, batch_size, images // batch_size))
    return out


def train(cfg, dataset_for_size, images=600000, passes=None, max_updates=None, **model_kw):
    """dataset_for_size(size) -> dataset object with .train / .test (the reference builds a TextDataset per stage from
    cfg.MODEL.SIZES, :57-63).  passes: optional subset of schedule indices."""
    models = []
    for i, (stage, prev, t, batch_size, max_iters) in enumerate(schedule(images)):
        if passes is not None and i not in passes:
            continue
        write = os.path.join(cfg.CHECKPOINT_DIR, 'stage%d/' % stage)
        read = os.path.join(cfg.CHECKPOINT_DIR, 'stage%d/' % prev)
        tag = 'stage_t%d/' % stage if t else 'stage%d/' % stage
        for p in (write, read):
            if not os.path.exists(p):
                os.makedirs(p)
        pggan = PGGAN(batch_size=batch_size, steps=max_iters, check_dir_write=write, check_dir_read=read,
                      dataset=dataset_for_size(cfg.MODEL.SIZES[stage - 1]), sample_path=os.path.join(cfg.SAMPLE_DIR, tag),
                      log_dir=os.path.join(cfg.LOGS_DIR, tag), stage=stage, trans=t, **model_kw)
        pggan.train(max_updates=max_updates)
        models.append(pggan)
    return models


if __name__ == "__main__":
    # python -m t2i_b200.models.pggan.train_pggan --cfg <yml> [--passes 0,1] [--max-updates N]   (synthetic batches: the
    # reference's pickled datasets, preprocess/dataset.py, are out of scope)
    import argparse

    from ...utils.config import config_from_yaml
    from ..wgancls.trainer import SyntheticTextDataset

    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", required=True, help="YAML with CHECKPOINT_DIR, SAMPLE_DIR, LOGS_DIR, MODEL.SIZES (models/pggan/cfg/*.yml)")
    ap.add_argument("--passes", default=None, help="comma-separated indices into the 15-entry schedule")
    ap.add_argument("--max-updates", type=int, default=None)
    a = ap.parse_args()
    cfg = config_from_yaml(a.cfg)
    train(cfg, lambda size: SyntheticTextDataset(embed_dim=cfg.MODEL.EMBED_DIM, image_size=size),
          passes=None if a.passes is None else [int(x) for x in a.passes.split(",")], max_updates=a.max_updates)
