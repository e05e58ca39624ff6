"""utils/saver.py of the reference (:6-25): save / load with the step recovered from the checkpoint
file name by the same regex.

The container is an .npz keyed by the reference's TF variable names in TF layouts (``model.get_variables()``)
plus this implementation's own optimizer-state keys (``__adam__/m/<var>``, ``__adam__/v/<var>``, step counters).
It is NOT a TF checkpoint: ``global_step`` is stored under that name (TF: ``Variable``), the Adam slots use the
keys above (TF: ``<var>/Adam``, ``<var>/Adam_1``), and ``set_variables`` wants every variable of the saved scopes.

What is kept of ``tf.train.Saver`` (the reference's ``saver.save(sess, dir, global_step=step)`` /
``tf.train.get_checkpoint_state(dir)``) is its notion of "latest", which is RECENCY, not the step number:
  * every save() rewrites the state file ``<dir>/checkpoint`` (TF's text format: ``model_checkpoint_path`` = the file
    just written, ``all_model_checkpoint_paths`` = the files this saver still keeps);
  * load() restores ``model_checkpoint_path`` — the most recently WRITTEN checkpoint, even when an older pass left files
    with larger step numbers in the same directory (the PGGAN schedule: a stabilisation pass restarts idx at 1 in the
    directory the transition pass filled, models/pggan/train_pggan.py:17-69);
  * ``max_to_keep`` prunes only among the files THIS saver object wrote, oldest first (tf.train.Saver keeps its own
    ``last_checkpoints`` list and never deletes files of earlier processes / other savers).
Directories without a state file (written by an earlier version) fall back to the newest file by modification time.

Multi-process (``torch.distributed`` initialised): rank 0 alone writes and prunes, every rank passes a barrier
afterwards, so that replicas never race on the same path.
"""
import os
import re

import numpy as np

STATE_FILE = "checkpoint"
_PAT = r"%s-\d+\.npz$"


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except Exception:       # torch absent from a pure-numpy caller: single process
        pass
    return None


def _write_state(checkpoint_dir, latest, kept):
    tmp = os.path.join(checkpoint_dir, STATE_FILE + ".tmp")
    with open(tmp, "w") as f:
        f.write('model_checkpoint_path: "%s"\n' % latest)
        for k in kept:
            f.write('all_model_checkpoint_paths: "%s"\n' % k)
    os.replace(tmp, os.path.join(checkpoint_dir, STATE_FILE))


def _read_state(checkpoint_dir):
    path = os.path.join(checkpoint_dir, STATE_FILE)
    if not os.path.isfile(path):
        return None
    with open(path) as f:
        for line in f:
            m = re.match(r'model_checkpoint_path:\s*"(.*)"\s*$', line)
            if m:
                return m.group(1)
    return None


def save(model, checkpoint_dir, step, max_to_keep=3, prefix="wgancls"):
    """Write ``<prefix>-<step>.npz``; returns its path.  ``model`` is anything with get_variables() /
    get_optimizer_state() (a model, or a variable-subset saver object); it also carries the list of files it wrote."""
    dist = _dist()
    path = os.path.join(checkpoint_dir, "%s-%d.npz" % (prefix, step))
    if dist is None or dist.get_rank() == 0:
        if not os.path.exists(checkpoint_dir):
            os.makedirs(checkpoint_dir)
        variables = {k: np.asarray(v) for k, v in model.get_variables().items()}
        opt = model.get_optimizer_state()
        variables.update({"__adam__/" + k: np.asarray(v) for k, v in opt.items()})
        tmp = path + ".tmp.npz"
        np.savez(tmp, **variables)
        os.replace(tmp, path)                     # never leave a torn file under the final name
        name = os.path.basename(path)
        mine = model.__dict__.setdefault("_saver_last_checkpoints", {}).setdefault(os.path.abspath(checkpoint_dir), [])
        if name in mine:
            mine.remove(name)                     # re-written: it becomes the most recent
        mine.append(name)
        while max_to_keep and len(mine) > max_to_keep:
            old = mine.pop(0)
            try:
                os.remove(os.path.join(checkpoint_dir, old))
            except FileNotFoundError:
                pass
        _write_state(checkpoint_dir, name, mine)
    else:
        model.get_variables()                     # the same device synchronisations on every rank
    if dist is not None:
        dist.barrier()
    return path


def latest_checkpoint(checkpoint_dir, prefix="wgancls"):
    """File name of the most recently written checkpoint in the directory, or None (tf.train.get_checkpoint_state)."""
    if not os.path.isdir(checkpoint_dir):
        return None
    name = _read_state(checkpoint_dir)
    if name is not None and os.path.isfile(os.path.join(checkpoint_dir, name)):
        return name
    names = [f for f in os.listdir(checkpoint_dir) if re.match(_PAT % re.escape(prefix), f)]
    if not names:
        return None
    return max(names, key=lambda f: (os.path.getmtime(os.path.join(checkpoint_dir, f)), f))


def load(model, checkpoint_dir, prefix="wgancls"):
    print(" [*] Reading checkpoints from %s..." % checkpoint_dir)
    ckpt_name = latest_checkpoint(checkpoint_dir, prefix)
    if ckpt_name:
        z = np.load(os.path.join(checkpoint_dir, ckpt_name))
        model.set_variables({k: z[k] for k in z.files if not k.startswith("__adam__/")})
        model.set_optimizer_state({k[len("__adam__/"):]: z[k] for k in z.files if k.startswith("__adam__/")})
        counter = int(next(re.finditer(r"(\d+)(?!.*\d)", ckpt_name)).group(0))
        print(" [*] Success to read {}".format(ckpt_name))
        return True, counter
    print(" [*] Failed to find checkpoints")
    return False, 0
