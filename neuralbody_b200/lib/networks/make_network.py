"""Plugin factory, same contract as the reference's lib/networks/make_network.py:5-9:
the class is picked by the file path in `cfg.network_path` (`imp.load_source` upstream;
`imp` is gone in Python >= 3.12, so importlib does the same job)."""
import importlib.util
import sys


def load_source(module_name, path):
    if module_name in sys.modules and getattr(sys.modules[module_name], "__file__", None) == path:
        return sys.modules[module_name]
    spec = importlib.util.spec_from_file_location(module_name, path)
    module = importlib.util.module_from_spec(spec)
    sys.modules[module_name] = module
    spec.loader.exec_module(module)
    return module


def make_network(cfg):
    plugin = load_source(cfg.network_module, cfg.network_path)
    return plugin.Network()
