"""ConfigMixin / register_to_config (diffusers 0.17.1 semantics: ctor kwargs recorded into `self.config`)."""
import functools
import inspect


class FrozenDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        raise AttributeError("config is frozen")


class ConfigMixin:
    config_name = "config.json"

    def register_to_config(self, **kwargs):
        kwargs.pop("kwargs", None)
        prev = dict(getattr(self, "_internal_dict", {}))
        prev.update(kwargs)
        object.__setattr__(self, "_internal_dict", FrozenDict(prev))

    @property
    def config(self):
        return self._internal_dict


def register_to_config(init):
    @functools.wraps(init)
    def inner_init(self, *args, **kwargs):
        init_kwargs = {k: v for k, v in kwargs.items() if not k.startswith("_")}
        sig = inspect.signature(init)
        params = {n: p.default for i, (n, p) in enumerate(sig.parameters.items()) if i > 0}
        new_kwargs = {}
        for a, name in zip(args, params.keys()):
            new_kwargs[name] = a
        new_kwargs.update({k: init_kwargs.get(k, d) for k, d in params.items() if k not in new_kwargs})
        new_kwargs = {**new_kwargs, **init_kwargs}
        init(self, *args, **init_kwargs)
        # the outermost (sub)class registers last and therefore wins, like diffusers
        ConfigMixin.register_to_config(self, **new_kwargs)

    return inner_init
