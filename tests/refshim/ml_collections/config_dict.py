"""ml_collections.config_dict.ConfigDict: attribute + item access, nested (what config.py builds)."""


class ConfigDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items(): self[k] = v
    def __setitem__(self, k, v): super().__setitem__(k, ConfigDict(v) if isinstance(v, dict) and not isinstance(v, ConfigDict) else v)
    def __getattr__(self, k):
        try: return self[k]
        except KeyError as e: raise AttributeError(k) from e
    def __setattr__(self, k, v): self[k] = v
