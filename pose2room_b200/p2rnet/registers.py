"""Name -> class registries with the reference's interface (net_utils/registry.py:5-47,
models/registers.py:6-8): `REG.get(key, alter_key=None)`, `@REG.register_module`, duplicate names raise
KeyError.  `install_into(reference_registries)` publishes the B200 classes under the reference's own
registries so `METHODS.get('P2RNet')(cfg)` in net_utils/utils.py:247 builds this model."""
import inspect


class Registry(object):
    def __init__(self, name):
        self._name = name
        self._module_dict = {}

    def __repr__(self):
        return "%s(name=%s, items=%s)" % (type(self).__name__, self._name, list(self._module_dict))

    @property
    def name(self):
        return self._name

    @property
    def module_dict(self):
        return self._module_dict

    def get(self, key, alter_key=None):
        if key in self._module_dict:
            return self._module_dict[key]
        return self._module_dict.get(alter_key, None)

    def register_module(self, cls):
        if not inspect.isclass(cls):
            raise TypeError("module must be a class, but got %s" % type(cls))
        if cls.__name__ in self._module_dict:
            raise KeyError("%s is already registered in %s" % (cls.__name__, self._name))
        self._module_dict[cls.__name__] = cls
        return cls


METHODS = Registry("method")
MODULES = Registry("module")
LOSSES = Registry("loss")


def install_into(ref_methods, ref_modules, ref_losses):
    """Overwrite the reference registries' entries with the B200 classes (same names)."""
    for src, dst in [(METHODS, ref_methods), (MODULES, ref_modules), (LOSSES, ref_losses)]:
        for name, cls in src.module_dict.items():
            dst.module_dict[name] = cls
