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
    With include_nets=True, `mmnas.model.full_vqa` / `hygr_vqa` also resolve to this package's nets."""
    from .model import modules, mixed, nets
    from .utils import ops_adapter
    root = sys.modules.get('mmnas')
    if root is None or getattr(root, '__mmnas_b200__', False) is False:
        root = types.ModuleType('mmnas')
        root.__path__ = []
        root.__mmnas_b200__ = True
        sys.modules['mmnas'] = root
    for pkg in ('mmnas.model', 'mmnas.utils'):
        if pkg not in sys.modules or not getattr(sys.modules[pkg], '__mmnas_b200__', False):
            m = types.ModuleType(pkg)
            m.__path__ = []
            m.__mmnas_b200__ = True
            sys.modules[pkg] = m
            setattr(root, pkg.split('.')[1], m)
    sys.modules['mmnas.model.modules'] = modules
    sys.modules['mmnas.model.mixed'] = mixed
    sys.modules['mmnas.utils.ops_adapter'] = ops_adapter
    sys.modules['mmnas.model'].modules = modules
    sys.modules['mmnas.model'].mixed = mixed
    sys.modules['mmnas.utils'].ops_adapter = ops_adapter
    if include_nets:
        for name in ('full_vqa', 'full_vgd', 'full_itm', 'hygr_vqa', 'hygr_vgd', 'hygr_itm'):
            sys.modules['mmnas.model.' + name] = nets
            setattr(sys.modules['mmnas.model'], name, nets)
