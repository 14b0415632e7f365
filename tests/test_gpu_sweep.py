"""BASELINE configs[4]: cell-size sweep 0.05 .. 1.0 m with the occupancy-skew variant
(20 % of the points in 0.1 % of the area), on a subsample the oracle finishes in seconds."""
import pytest

from grid_ndt_b200 import synthetic
from grid_ndt_b200._abi import default_params

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("grid_len", [0.05, 0.1, 0.2, 0.5, 1.0])
@pytest.mark.parametrize("skew", [False, True])
def test_cell_size_sweep_parity(grid_len, skew):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test on a box without CUDA (no fallback exists)")
    from tests import parity
    cloud = synthetic.cfg5(1_500_000, extent=60.0, skew=skew)
    rep = parity.run_case(cloud, default_params(grid_len, 0.1, 0.08), "slope")
    assert rep["ok"], rep["fail"]
    assert rep["exact_count_mismatch"] == 0 and rep["label_mismatch"] == rep["label_mismatch_agreeing_with_truth64"] + rep["label_mismatch_threshold_adjacent"]
