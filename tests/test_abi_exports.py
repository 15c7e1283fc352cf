"""CPU: the C-ABI library loads, exports every symbol include/sage_splat.h declares, and fails
loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    text = (ROOT / "include" / "sage_splat.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sage_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(sage):
    from sage3d_official_b200 import _lib
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in sage_splat.h but not exported"
    assert sorted(_lib.EXPORTS) == names, "EXPORTS list out of sync with the header"
    assert L.sage_abi_version() == 2


def test_struct_layouts_match_header(sage):
    from sage3d_official_b200 import _lib
    assert C.sizeof(_lib.CCamera) == 4 * (9 + 3 + 3 + 4) + 4 * 2 + 4 * 2
    assert C.sizeof(_lib.CFrameStats) == 8 * 6 + 4 * 2
    assert C.sizeof(_lib.CSplatRecord) == 4 * 14


def test_fails_loudly_without_gpu(sage):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cloud, _ = sage.synthetic_room(0, 32)
    with pytest.raises(sage.SageError, match="no CPU fallback"):
        sage.GaussianScene(cloud)
    with pytest.raises(sage.SageError, match="no CPU fallback"):
        sage.ObstacleGrid(np.zeros((4, 4), np.uint8), 0, 1, 0, 1, 0.05)
    with pytest.raises(sage.SageError, match="no CPU fallback"):
        sage.CollisionBodies(np.zeros((1, 3)), np.zeros((1, 3)), np.array([[1.0, 0, 0, 0]]))


def test_product_never_imports_oracle():
    pkg = ROOT / "sage-3d_official_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        txt = f.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt and "sage_oracle" not in txt, f
