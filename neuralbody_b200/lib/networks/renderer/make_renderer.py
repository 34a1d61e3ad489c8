"""Plugin factory, same contract as the reference's lib/networks/renderer/make_renderer.py:5-9."""
from ..make_network import load_source


def make_renderer(cfg, network):
    module = cfg.renderer_module
    path = cfg.renderer_path
    renderer = load_source(module, path).Renderer(network)
    return renderer
