"""Genotypes (searched architectures) in the reference's arch/*.json format.

An arch file is `{"epoch<N>": {"enc": [[op], ...], "dec": [[op], ...]}}` (written by search_vqa.py:363-386, read
by train_vqa.py:185).  `load_arch` consumes such a file unchanged; `dump_arch` writes one.  The four
architectures the reference ships are kept here as compact codes (S = self_att_64, R = rel_self_att_64,
G = guided_att_64, F = feed_forward; one letter per node) and expanded on demand, so the repository carries
no copy of the reference's data files; tests/test_genotypes.py checks the expansion against
/root/reference/arch/*.json when that tree is mounted.
"""
import json

CODE = {'S': 'self_att_64', 'R': 'rel_self_att_64', 'G': 'guided_att_64', 'F': 'feed_forward'}
LETTER = {v: k for k, v in CODE.items()}

SHIPPED = {
    'mcan':      ('SFSFSFSFSFSF', 'SGFSGFSGFSGFSGFSGF'),
    'mmnas_vqa': ('SSSSFFFFSFFF', 'GGFFGFRGFGRFRSFRGF'),
    'mmnas_vgd': ('SFFSFFFFFFFS', 'GGGGGGFGRRGFRGGRGR'),
    'mmnas_itm': ('SSFFFSFSFFFF', 'SGGRSGRGGGGFGGRSGR'),
}


def expand(enc_code, dec_code):
    return {'enc': [[CODE[c]] for c in enc_code], 'dec': [[CODE[c]] for c in dec_code]}


def compact(genotype):
    return tuple(''.join(LETTER[node[0]] for node in genotype[k]) for k in ('enc', 'dec'))


def shipped(name):
    return expand(*SHIPPED[name])


def load_arch(path, epoch=0):
    """Same access pattern as train_vqa.py:185: json.load(path)['epoch' + str(GENO_EPOCH)]."""
    with open(path) as f:
        return json.load(f)['epoch' + str(epoch)]


def dump_arch(path, genotype, epoch=0, merge=True):
    data = {}
    if merge:
        try:
            with open(path) as f:
                data = json.load(f)
        except (OSError, ValueError):
            data = {}
    data['epoch' + str(epoch)] = genotype
    with open(path, 'w') as f:
        json.dump(data, f)


def count_ops(genotype):
    out = {}
    for k in ('enc', 'dec'):
        for node in genotype[k]:
            for n in node:
                out[(k, n)] = out.get((k, n), 0) + 1
    return out
