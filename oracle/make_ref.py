"""oracle/make_ref.py -- TEST INFRASTRUCTURE, not product code.

Builds `oracle/_ref/l2hmc`: a writable, git-ignored copy of the reference's
hot-path Python modules (the reference is 100 % Python, so "building" it is a
copy plus the one-dataclass patch Python >= 3.11 needs).  It exists so that

  * `oracle/make_golden.py` can run the reference's own PyTorch path here and
    freeze its outputs into `tests/golden/`, and
  * `bench.py --impl reference` / `cpu_baseline` can time the reference's own
    CPU path on the GPU box's host cores (`oracle/_ref` travels with gpurun,
    `/root/reference` does not).

Only the modules the hot path imports are copied:
configs, group/, lattice/, dynamics/pytorch, network/, loss/.  Nothing from
`oracle/_ref` is ever committed (see .gitignore) and nothing in the product
package imports it.

Why a copy and not an in-place import: `l2hmc/configs.py:42-46` creates
directories next to the source at import time and `/root/reference` is
read-only; `configs.py:298-302` uses a mutable dataclass default that Python
3.12 rejects.
"""
from __future__ import annotations

import os
import re
import shutil
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
DEFAULT_SRC = Path(os.environ.get('L2HMC_REFERENCE', '/root/reference')) / 'src' / 'l2hmc'
# one level down so the reference's PROJECT_DIR (configs.py:33) resolves to oracle/_ref
DST = HERE / '_ref' / 'src' / 'l2hmc'

# sub-trees of src/l2hmc that the hot path touches
KEEP = [
    '__init__.py', '__about__.py', 'configs.py',
    'group', 'lattice', 'dynamics', 'network', 'loss',
]
SKIP_DIRS = {'tensorflow', '__pycache__', 'numpy'}


def _ignore(_dir: str, names: list[str]) -> list[str]:
    return [n for n in names if n in SKIP_DIRS or n.endswith('.pyc')]


def build(src: Path = DEFAULT_SRC, dst: Path = DST, force: bool = False) -> bool:
    """Returns True if `dst` is usable afterwards."""
    if dst.exists() and not force:
        return True
    if not src.exists():
        return False
    if dst.exists():
        shutil.rmtree(dst)
    dst.mkdir(parents=True)
    for name in KEEP:
        s = src / name
        if not s.exists():
            continue
        if s.is_dir():
            shutil.copytree(s, dst / name, ignore=_ignore)
        else:
            shutil.copy2(s, dst / name)
    # lattice/u1/numpy is an independent numpy cross-check; keep it
    s = src / 'lattice' / 'u1' / 'numpy'
    if s.exists():
        shutil.copytree(s, dst / 'lattice' / 'u1' / 'numpy', ignore=_ignore)
    # --- the one patch (configs.py:298-302): mutable dataclass defaults -----
    cfg = dst / 'configs.py'
    txt = cfg.read_text()
    txt, n = re.subn(
        r'^(\s+)(x|v): NetWeight = NetWeight\(1\., 1\., 1\.\)\s*$',
        r'\1\2: NetWeight = field(default_factory=lambda: NetWeight(1., 1., 1.))',
        txt, flags=re.M)
    assert n == 2, f'expected to patch 2 dataclass defaults, patched {n}'
    cfg.write_text(txt)
    return True


if __name__ == '__main__':
    ok = build(force='--force' in sys.argv)
    print(f'oracle/_ref: {"ready" if ok else "reference sources not found"} -> {DST}')
