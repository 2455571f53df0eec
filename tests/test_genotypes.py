import json
import os

import pytest
import torch

from mmnas_b200 import genotypes as G
from mmnas_b200.data.synthetic import box_geometry
from tests.util import load_golden

REF_ARCH = '/root/reference/arch'


def test_roundtrip(tmp_path):
    g = G.shipped('mmnas_vqa')
    assert len(g['enc']) == 12 and len(g['dec']) == 18
    assert G.compact(g) == G.SHIPPED['mmnas_vqa']
    p = tmp_path / 'a.json'
    G.dump_arch(str(p), g, epoch=3)
    G.dump_arch(str(p), G.shipped('mcan'), epoch=4)
    assert G.load_arch(str(p), 3) == g and G.load_arch(str(p), 4) == G.shipped('mcan')
    assert G.count_ops(g)[('dec', 'rel_self_att_64')] == 4


@pytest.mark.skipif(not os.path.isdir(REF_ARCH), reason='reference not mounted')
@pytest.mark.parametrize('name', sorted(G.SHIPPED))
def test_shipped_codes_equal_reference_arch_files(name):
    ref = json.load(open(os.path.join(REF_ARCH, name + '.json')))
    assert ref == {'epoch0': G.shipped(name)}
    assert G.load_arch(os.path.join(REF_ARCH, name + '.json'), 0) == G.shipped(name)


def test_box_geometry_matches_reference_golden():
    z = load_golden('geometry.npz')
    assert torch.allclose(box_geometry(z['boxes']), z['rel'], rtol=0, atol=0)
