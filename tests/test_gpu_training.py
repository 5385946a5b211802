"""GPU smoke of the procedural multi-traversal training loop (stand-in for BASELINE configs 3-4): the public API
(spherical_harmonics -> rasterization -> backward, absgrad statistics) drives an optimiser and PSNR improves."""
import importlib.util
import os

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_multitraversal_training_improves_psnr(cuda_device):
    spec = importlib.util.spec_from_file_location("train_mt", os.path.join(ROOT, "examples", "train_multitraversal.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    first, final = mod.main(["--n-gauss", "20000", "--traversals", "2", "--width", "320", "--height", "192", "--iters", "80",
                             "--log-every", "40"])
    assert set(first) == set(final) == {0, 1}
    for t in first:
        assert final[t] > first[t] + 1.0, (first, final)
