"""`torchdrug.core` stand-in: Registry / Configurable as used by reference model.py:17-18."""
import inspect


class Registry(object):
    """`@R.register("models.Name")` decorator + lookup table."""

    table = {}

    @classmethod
    def register(cls, name):
        def wrapper(obj):
            cls.table[name] = obj
            short = name.split(".")[-1]
            cls.table.setdefault(short, obj)
            return obj
        return wrapper

    @classmethod
    def get(cls, name):
        if name not in cls.table:
            raise KeyError("Can't find `%s` in the registry" % name)
        return cls.table[name]

    @classmethod
    def search(cls, name):
        return cls.get(name)


class Configurable(object):
    """Config-dict <-> object factory (reference run_full.py:49-51 builds tasks this way)."""

    @classmethod
    def load_config_dict(cls, config):
        config = dict(config)
        name = config.pop("class")
        target = Registry.get(name) if cls is Configurable or name != cls.__name__ else cls
        kwargs = {}
        for key, value in config.items():
            if isinstance(value, dict) and "class" in value:
                value = Configurable.load_config_dict(value)
            kwargs[key] = value
        return target(**kwargs)

    def config_dict(self):
        signature = inspect.signature(type(self).__init__)
        config = {"class": type(self).__name__}
        for key in list(signature.parameters)[1:]:
            if hasattr(self, key):
                config[key] = getattr(self, key)
        return config


def make_configurable(cls, module=None, ignore_args=()):
    return type(cls.__name__, (cls, Configurable), {})


class Meter(object):
    """Placeholder: the training engine is out of scope."""

    def __init__(self, *args, **kwargs):
        self.records = {}

    def update(self, record):
        for key, value in record.items():
            self.records.setdefault(key, []).append(float(value))


class Engine(object):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("torchdrug.core.Engine is outside the rspmm hot path (SURVEY.md 2.1 #5)")
