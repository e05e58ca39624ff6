"""utils/saver.py of the reference (:6-25): save / load with the step recovered from the checkpoint
file name by the same regex.  The container is an .npz of the reference's TF variable names and
layouts (model.get_variables()), so a tensor-by-tensor dump of a TF checkpoint loads directly."""
import os
import re

import numpy as np


def save(model, checkpoint_dir, step, max_to_keep=3):
    if not os.path.exists(checkpoint_dir):
        os.makedirs(checkpoint_dir)
    variables = {k: np.asarray(v) for k, v in model.get_variables().items()}
    opt = model.get_optimizer_state()
    variables.update({"__adam__/" + k: np.asarray(v) for k, v in opt.items()})
    path = os.path.join(checkpoint_dir, "wgancls-%d.npz" % step)
    np.savez(path, **variables)
    kept = sorted((f for f in os.listdir(checkpoint_dir) if re.match(r"wgancls-\d+\.npz$", f)),
                  key=lambda f: int(re.findall(r"\d+", f)[-1]))
    for f in kept[:-max_to_keep]:
        os.remove(os.path.join(checkpoint_dir, f))
    return path


def load(model, checkpoint_dir):
    print(" [*] Reading checkpoints from %s..." % checkpoint_dir)
    names = [f for f in os.listdir(checkpoint_dir)] if os.path.isdir(checkpoint_dir) else []
    names = sorted((f for f in names if re.match(r"wgancls-\d+\.npz$", f)), key=lambda f: int(re.findall(r"\d+", f)[-1]))
    if names:
        ckpt_name = names[-1]
        z = np.load(os.path.join(checkpoint_dir, ckpt_name))
        model.set_variables({k: z[k] for k in z.files if not k.startswith("__adam__/")})
        model.set_optimizer_state({k[len("__adam__/"):]: z[k] for k in z.files if k.startswith("__adam__/")})
        counter = int(next(re.finditer(r"(\d+)(?!.*\d)", ckpt_name)).group(0))
        print(" [*] Success to read {}".format(ckpt_name))
        return True, counter
    print(" [*] Failed to find checkpoints")
    return False, 0
