"""The in-library multi-device entry points (-m gpu): one host process, ctypes only, NCCL communicators created inside the
library — on however many devices the box has (a 1-GPU box runs the same code with a single-rank communicator; the 2- and
8-GPU runs are recorded under profiles/)."""
import pytest

pytestmark = pytest.mark.gpu


def test_multi_device_entry_points_match_single_device(gpu):
    from tools import multi_check
    res = multi_check.run(None, n_solve=1500, n_cluster=20000, verbose=False)
    print("multi-device check:", res)
    assert res["n_devices"] >= 1 and res["nccl_version"] > 0
    assert res["solve_equal"], "mcbb_solve_features_multi differs from mcbb_solve_features"
    assert res["dbscan_equal"], "mcbb_dbscan_features_multi differs from mcbb_dbscan_features"
    assert res["pipeline_equal"], "mcbb_solve_cluster_multi differs from solve -> sort -> dbscan"
