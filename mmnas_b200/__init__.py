"""mmnas_b200 — B200-native operator hot path of MMnas (SA / GA / RSA / FFN blocks + supernet mixed-op).

    import mmnas_b200
    mmnas_b200.set_precision('bf16')          # or 'fp32'
    from mmnas_b200.utils.ops_adapter import OpsAdapter
    op = OpsAdapter().OPS['rel_self_att_64'](cfg, norm=True, residual=True).cuda()

`install_as_mmnas()` registers this package's modules under the reference's import paths
(`mmnas.model.modules`, `mmnas.model.mixed`, `mmnas.utils.ops_adapter`), so the reference's own
full_*.py / hygr_*.py / train_*.py / search_*.py pick up the CUDA operators unchanged.
"""
import sys
import types

from .runtime import set_precision, get_precision, precision, manual_seed, advance  # noqa: F401

__version__ = '0.1.0'


def install_as_mmnas(include_nets=False):
    """Alias the drop-in modules under the reference's package name.  Call before importing reference code.

    If the reference package `mmnas` is importable (its checkout is on sys.path) it is kept, and only
    `mmnas.model.modules`, `mmnas.model.mixed` and `mmnas.utils.ops_adapter` are replaced, so the reference's own
    full_*.py / hygr_*.py / train / search code runs on the CUDA operators.  Without the reference, a synthetic
    `mmnas` package is created.  With include_nets=True, `mmnas.model.full_*` / `hygr_*` also resolve to this
    package's nets."""
    import importlib
    from .model import modules, mixed, nets
    from .utils import ops_adapter
    pkgs = {}
    for pkg in ('mmnas', 'mmnas.model', 'mmnas.utils'):
        try:
            pkgs[pkg] = importlib.import_module(pkg)
        except ImportError:
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m
            pkgs[pkg] = m
            if '.' in pkg:
                setattr(pkgs['mmnas'], pkg.split('.')[1], m)
    for full, mod in (('mmnas.model.modules', modules), ('mmnas.model.mixed', mixed),
                      ('mmnas.utils.ops_adapter', ops_adapter)):
        sys.modules[full] = mod
        parent, leaf = full.rsplit('.', 1)
        setattr(pkgs[parent], leaf, mod)
    if include_nets:
        for name in ('full_vqa', 'full_vgd', 'full_itm', 'hygr_vqa', 'hygr_vgd', 'hygr_itm'):
            sys.modules['mmnas.model.' + name] = nets
            setattr(pkgs['mmnas.model'], name, nets)
