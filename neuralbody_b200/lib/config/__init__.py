from .config import cfg, make_cfg, CfgNode, get_active_cfg  # noqa: F401
