"""Populate baseline/_ref/ with an UNMODIFIED, byte-identical copy of the reference's model package.

    python scripts/install_reference.py [/root/reference]

The reference (MILVLG/mmnas, Apache-2.0) is pure Python without packaging metadata, so `pip install --target
baseline/_ref /root/reference` has nothing to install (its only setup.py builds the out-of-scope Cython IoU).  What the
GPU box needs of it is the model package the hot path plugs into: mmnas/model/*.py (the callers full_*.py / hygr_*.py
and the reference operators modules.py / mixed.py), mmnas/utils/{ops_adapter,optimizer,itm_loss}.py and arch/*.json.
They are copied file by file; baseline/_ref/MANIFEST.json records the sha256 of every file so tests can prove the copy
is untouched.  baseline/_ref/ is git-ignored (never product source, never in history) but travels with the gpurun
snapshot; it is used by
  * tests/test_gpu_reference_dropin.py — the reference's own Net_Full / Net_Search running on this library's
    operators (mmnas_b200.install_as_mmnas()) against the untouched reference on the same GPU;
  * bench.py --impl reference — the reference train step on the host cores (cpu_baseline.kind = "reference").
Nothing under mmnas_b200/ imports it.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, 'baseline', '_ref')
FILES = ['mmnas/model/modules.py', 'mmnas/model/mixed.py', 'mmnas/model/full_vqa.py', 'mmnas/model/full_vgd.py',
         'mmnas/model/full_itm.py', 'mmnas/model/hygr_vqa.py', 'mmnas/model/hygr_vgd.py', 'mmnas/model/hygr_itm.py',
         'mmnas/utils/ops_adapter.py', 'mmnas/utils/optimizer.py', 'mmnas/utils/itm_loss.py',
         'arch/mcan.json', 'arch/mmnas_vqa.json', 'arch/mmnas_vgd.json', 'arch/mmnas_itm.json', 'LICENSE']


def sha(path):
    return hashlib.sha256(open(path, 'rb').read()).hexdigest()


def install(src='/root/reference', verbose=True):
    """Returns the manifest, or None when the reference checkout is not present (the GPU box: the prebuilt copy that
    travelled with the snapshot is used as is)."""
    if not os.path.isdir(os.path.join(src, 'mmnas', 'model')):
        return None
    manifest = {}
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(DEST, rel)
        if not os.path.exists(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not os.path.exists(d) or sha(d) != sha(s):
            shutil.copyfile(s, d)
        manifest[rel] = sha(d)
    json.dump({'source': src, 'files': manifest}, open(os.path.join(DEST, 'MANIFEST.json'), 'w'), indent=1)
    if verbose:
        print('baseline/_ref: %d files from %s' % (len(manifest), src))
    return manifest


def verify():
    """True iff baseline/_ref exists and every file still has the recorded sha256."""
    mpath = os.path.join(DEST, 'MANIFEST.json')
    if not os.path.exists(mpath):
        return False
    files = json.load(open(mpath))['files']
    return all(os.path.exists(os.path.join(DEST, rel)) and sha(os.path.join(DEST, rel)) == h for rel, h in files.items())


if __name__ == '__main__':
    m = install(sys.argv[1] if len(sys.argv) > 1 else '/root/reference')
    sys.exit(0 if (m is not None and verify()) else 1)
