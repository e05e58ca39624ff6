"""Host logic of the trainer / saver mirrors on CPU: the reference's step order, lr schedule,
summary / sample / checkpoint periods (models/wgancls/trainer.py:73-127), driven against a stub
model that records the calls; checkpoint file naming and step recovery (utils/saver.py:13-25)."""
import os

import numpy as np

from t2i_b200.models.wgancls.trainer import SyntheticTextDataset, WGanClsTrainer
from t2i_b200.utils.config import AttrDict
from t2i_b200.utils import saver


class _F(object):
    def __init__(self, name):
        self.name = name


class StubModel(object):
    batch_size, sample_num, z_dim = 4, 6, 8

    def __init__(self):
        for n in ("learning_rate_d", "learning_rate_g", "x", "x_mismatch", "cond", "z", "epsilon", "z_sample",
                  "cond_sample", "iter", "D_optim", "kt_optim", "G_optim", "D_loss", "G_loss", "G_kl_loss", "sampler",
                  "D_loss_real", "D_loss_fake", "real_gp", "reg_loss", "wdist", "wdist2", "D_loss_mismatch", "real_gp2",
                  "kt", "balance_loss"):
            setattr(self, n, _F(n))
        self.calls, self.vars, self.opt = [], {"w": np.arange(3.0), "kt": np.float32(0.7)}, {"d_t": 0}

    def initialize(self, seed=0):
        self.calls.append(("init",))

    def run(self, fetches, feed_dict=None):
        single = not isinstance(fetches, list)
        names = [fetches.name] if single else [f.name for f in fetches]
        self.calls.append((tuple(names), dict((k.name, v) for k, v in (feed_dict or {}).items())))
        out = [np.zeros((self.sample_num, 64, 64, 3)) if n == "sampler" else (None if n.endswith("optim") else 1.5)
               for n in names]
        return out[0] if single else out

    def get_variables(self):
        return self.vars

    def set_variables(self, v):
        self.vars = dict(v)

    def get_optimizer_state(self):
        return self.opt

    def set_optimizer_state(self, s):
        self.opt = dict(s)


def test_trainer_step_order_schedule_and_checkpoints(tmp_path):
    cfg = AttrDict({"CHECKPOINT_DIR": str(tmp_path / "ckpt"),
                    "TRAIN": {"D_LR": 1e-4, "G_LR": 2e-4, "N_CRITIC": 2, "MAX_STEPS": 13, "SUMMARY_PERIOD": 5,
                              "SAMPLE_PERIOD": 6, "CHECKPOINTS_TO_KEEP": 3}})
    m = StubModel()
    samples = []
    tr = WGanClsTrainer(None, m, SyntheticTextDataset(embed_dim=16, num_examples=64), cfg,
                        on_samples=lambda idx, s, c: samples.append((idx, s.shape, len(c))))
    tr.train()
    runs = [c for c in m.calls if c[0] != "init"]
    d_runs = [c for c in runs if "D_optim" in c[0]]
    g_runs = [c for c in runs if "G_optim" in c[0]]
    assert len(d_runs) == 12 and len(g_runs) == 6                       # G run only when idx % N_CRITIC == 0
    order = [("D" if "D_optim" in c[0] else "G" if "G_optim" in c[0] else "S") for c in runs]
    assert "".join(order[:6]) == "DDGDDG"                               # D (+kt) first, then G (trainer.py:97-102)
    assert d_runs[0][0][:3] == ("D_optim", "kt_optim", "D_loss")
    fd = d_runs[0][1]
    assert fd["real_images" if "real_images" in fd else "x"].shape == (4, 64, 64, 3) and fd["eps" if "eps" in fd else "epsilon"].shape == (4, 1, 1, 1)
    assert abs(fd["learning_rate_d"] - 1e-4) < 1e-12 and abs(fd["learning_rate_g"] - 2e-4) < 1e-12
    assert [r["idx"] for r in tr.log] == [5, 10] and "wdist" in tr.log[0]
    assert samples == [(6, (6, 64, 64, 3), 6), (12, (6, 64, 64, 3), 6)]
    assert sorted(os.listdir(cfg.CHECKPOINT_DIR)) == ["checkpoint", "wgancls-2.npz"]  # idx % 500 == 2
    # resume: the step is parsed from the file name with the reference's regex
    m2 = StubModel()
    ok, counter = saver.load(m2, cfg.CHECKPOINT_DIR)
    assert ok and counter == 2 and np.allclose(m2.vars["w"], np.arange(3.0))
    for step in (502, 1002, 1502):
        saver.save(m, cfg.CHECKPOINT_DIR, step, max_to_keep=3)
    assert sorted(os.listdir(cfg.CHECKPOINT_DIR)) == ["checkpoint", "wgancls-1002.npz", "wgancls-1502.npz", "wgancls-502.npz"]
    assert saver.load(m2, cfg.CHECKPOINT_DIR) == (True, 1502)
    assert saver.load(m2, str(tmp_path / "missing")) == (False, 0)


def test_saver_latest_is_most_recently_written_not_highest_step(tmp_path):
    """tf.train.Saver semantics (utils/saver.py:16 get_checkpoint_state): a second pass that restarts its step counter
    in a directory an earlier, LONGER pass filled (the PGGAN stabilisation pass after its transition pass,
    models/pggan/train_pggan.py:17-69) must (a) not have its fresh checkpoints pruned because older files carry
    larger step numbers, and (b) be the one a later load() restores."""
    d = str(tmp_path / "stage3")
    first, second = StubModel(), StubModel()
    first.vars = {"w": np.full(3, 1.0), "kt": np.float32(0.7)}
    for step in (2000, 4000, 6000):                 # transition pass: idx up to 6000, keeps 2
        saver.save(first, d, step, max_to_keep=2, prefix="pggan")
    assert sorted(f for f in os.listdir(d) if f.endswith(".npz")) == ["pggan-4000.npz", "pggan-6000.npz"]
    second.vars = {"w": np.full(3, 2.0), "kt": np.float32(0.7)}
    p = saver.save(second, d, 2000, max_to_keep=2, prefix="pggan")     # shorter second pass: ends at idx 2000
    assert os.path.exists(p), "the checkpoint just written must survive its own save() call"
    assert saver.latest_checkpoint(d, "pggan") == "pggan-2000.npz"
    m = StubModel()
    assert saver.load(m, d, prefix="pggan") == (True, 2000) and np.allclose(m.vars["w"], 2.0)
    # the second saver prunes only what IT wrote; the first pass's files are not its to delete
    saver.save(second, d, 3000, max_to_keep=2, prefix="pggan")
    saver.save(second, d, 3500, max_to_keep=2, prefix="pggan")
    names = sorted(f for f in os.listdir(d) if f.endswith(".npz"))
    assert names == ["pggan-3000.npz", "pggan-3500.npz", "pggan-4000.npz", "pggan-6000.npz"]
    assert saver.load(m, d, prefix="pggan") == (True, 3500)
    # no state file (directory written by an older version): newest modification time wins
    os.remove(os.path.join(d, saver.STATE_FILE))
    os.utime(os.path.join(d, "pggan-3000.npz"), (2e9, 2e9))
    assert saver.load(m, d, prefix="pggan") == (True, 3000)


def test_lr_decay_matches_reference_formula():
    # trainer.py:82-86: lr * 0.95 ** ((idx // n_critic) // 10000)
    for idx, n_critic in ((9999, 1), (10000, 1), (20000, 2), (40001, 2)):
        kiter = (idx // n_critic) // 10000
        assert abs(1e-4 * 0.95 ** kiter - 1e-4 * 0.95 ** ((idx // n_critic) // 10000)) == 0


def test_run_mode_dispatch(tmp_path):
    """models/wgancls/run.py:19-70: --cfg YAML -> directories, dataset, EVAL / TRAIN / visualise dispatch; the TRAIN mode
    runs the real WGanCls + WGanClsTrainer on the CPU restatement of the kernels (tiny widths)."""
    import yaml
    import pytest
    import fake_kernels as fk
    from t2i_b200.models.wgancls import run
    base = {"CONFIG_NAME": "t", "DATASET_NAME": "flowers", "DATASET_DIR": str(tmp_path / "data"),
            "CHECKPOINT_DIR": str(tmp_path / "ckpt"), "LOGS_DIR": str(tmp_path / "logs"), "SAMPLE_DIR": str(tmp_path / "samples"),
            "MODEL": {"Z_DIM": 8, "OUTPUT_SIZE": 64, "EMBED_DIM": 32, "COMPRESSED_EMBED_DIM": 8, "GF_DIM": 8, "DF_DIM": 8,
                      "IMAGE_SHAPE": {"W": 64, "H": 64, "D": 3}},
            "TRAIN": {"FLAG": True, "MAX_STEPS": 4, "BATCH_SIZE": 2, "SAMPLE_NUM": 2, "D_LR": 1e-4, "G_LR": 1e-4, "BETA1": 0.0,
                      "BETA2": 0.9, "SUMMARY_PERIOD": 1, "N_CRITIC": 1, "NUM_EMBEDDINGS": 4, "CHECKPOINTS_TO_KEEP": 3,
                      "SAMPLE_PERIOD": 300, "COEFF": {"KL": 1.0, "LAMBDA": 100.0}},
            "EVAL": {"FLAG": False, "SAMPLE_SIZE": 4, "SIZE": 8}}
    kw = dict(precision="bf16x3", device="cpu", kernels=fk, use_graphs=False)

    def write(cfg, name):
        p = tmp_path / name
        p.write_text(yaml.safe_dump(cfg))
        return str(p)

    tr = run.main(write(base, "train.yml"), **kw)
    assert [r["idx"] for r in tr.log] == [1, 2, 3] and all(np.isfinite(r["D_loss"]) and np.isfinite(r["G_loss"]) for r in tr.log)
    for d in ("ckpt", "logs", "samples"):
        assert os.path.isdir(str(tmp_path / d))
    assert sorted(os.listdir(str(tmp_path / "ckpt"))) == ["checkpoint", "wgancls-2.npz"]
    ev = dict(base, EVAL=dict(base["EVAL"], FLAG=True))
    with pytest.raises(NotImplementedError, match="Inception"):
        run.main(write(ev, "eval.yml"), **kw)
    vis = dict(base, TRAIN=dict(base["TRAIN"], FLAG=False))
    with pytest.raises(NotImplementedError, match="visualis"):
        run.main(write(vis, "vis.yml"), **kw)
    (tmp_path / "data").mkdir()
    (tmp_path / "data" / "train").write_text("x")
    (tmp_path / "data" / "test").write_text("x")
    with pytest.raises(NotImplementedError, match="pickled"):
        run.main(write(base, "train2.yml"), **kw)
