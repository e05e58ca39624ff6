"""utils/config.py of the reference (config_from_yaml, :5-7): YAML -> attribute-style dict.
The reference uses easydict + yaml.load; both are replaced by a tiny attr-dict and safe_load."""
import yaml


class AttrDict(dict):
    """dict whose keys are also attributes, recursively (what EasyDict gives the reference)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            v = AttrDict(v)
        elif isinstance(v, (list, tuple)):
            v = type(v)(AttrDict(x) if isinstance(x, dict) else x for x in v)
        super().__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    __setattr__ = __setitem__


def config_from_yaml(file_path):
    with open(file_path, "r") as f:
        return AttrDict(yaml.safe_load(f))
