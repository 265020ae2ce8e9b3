"""CPU-only: the padded structure-of-arrays loader against the per-epoch loader with
the reference's semantics, on the reference's own GNSS logs.  The logs live in the reference checkout, which
exists where the CPU suite runs and not on the GPU box: skipped when absent."""
import os

import numpy as np
import pytest

REF = "/root/reference/data"
LOGS = [
    REF + "/gnss_stationary/gnss_log_2020_02_05_09_14_15",
    REF + "/multi-receiver/rec1_gnss_log_2020_02_27_13_07_10",
    REF + "/gnss-multi-receiver/rec1/rec1_gnss_log_50y_moving_",
]


@pytest.mark.parametrize("prefix", LOGS)
def test_padded_loader_equals_the_per_epoch_loader(prefix):
    if not os.path.exists(prefix + "ranges.mat"):
        pytest.skip("reference data not present")
    from nlp_filter_b200.utils import data as du
    a = du.load_gnss_logs(prefix)
    b = du.load_gnss_logs_padded(prefix)
    T = len(a["pr"])
    assert b["pr"].shape[0] == T and b["sat_pos"].shape[:2] == b["pr"].shape and b["nsat"].shape == (T,)
    assert np.array_equal(np.asarray(list(a["t"]), dtype=float), b["t"])
    for k in range(T):
        n = a["pr"][k].shape[0]
        assert b["nsat"][k] == n
        assert np.array_equal(b["pr"][k, :n], a["pr"][k]) and np.all(b["pr"][k, n:] == 0.0)
        assert np.array_equal(b["sat_pos"][k, :n], a["sat_pos"][k]) and np.all(b["sat_pos"][k, n:] == 0.0)
        if "pr_rate" in a:
            assert np.array_equal(b["pr_rate"][k, :n], a["pr_rate"][k])
            assert np.array_equal(b["sat_vel"][k, :n], a["sat_vel"][k])
    wide = du.load_gnss_logs_padded(prefix, smax=b["pr"].shape[1] + 3)
    assert np.array_equal(wide["pr"][:, :b["pr"].shape[1]], b["pr"]) and np.all(wide["pr"][:, b["pr"].shape[1]:] == 0.0)
    with pytest.raises(ValueError):
        du.load_gnss_logs_padded(prefix, smax=max(int(b["nsat"].max()) - 1, 0))
