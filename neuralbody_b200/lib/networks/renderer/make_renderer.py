"""Renderer plugin factory: the contract of the reference's lib/networks/renderer/make_renderer.py:5-9 -- the yaml keys
`renderer_module` / `renderer_path` name a source file whose `Renderer(network)` is the plugin -- on importlib
(`imp`, which upstream uses, left the standard library in Python 3.12)."""
from ..make_network import load_source


def make_renderer(cfg, network):
    plugin = load_source(cfg.renderer_module, cfg.renderer_path)
    return plugin.Renderer(network)
