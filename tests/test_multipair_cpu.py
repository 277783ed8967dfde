"""Host-side logic of the multi-pair experiment driver: the pair cache's file names and order, the
convergence-report table's columns (reference run_hierarchical_optimizer3d_multipair.py:85-131) and the analysis file.
No GPU needed: the reports are built by hand."""
import os

import numpy as np
import pytest

from lsf_b200 import multipair, telemetry

# the reference's column list, written out (run_hierarchical_optimizer3d_multipair.py:88-106)
REFERENCE_LEVEL_COLUMNS = ["iter_count", "iter_lim_reached", "warp_delta_amt_ratio", "warp_delta_min", "warp_delta_max",
                           "warp_delta_mean", "warp_delta_std", "warp_delta_max_x", "warp_delta_max_y",
                           "warps_below_min_thresh", "warps_above_max_thresh", "diff_delta_min", "diff_delta_max",
                           "diff_delta_mean", "diff_delta_std", "diff_max_x", "diff_max_y"]


def make_report(iterations, limit_reached, seed):
    rng = np.random.default_rng(seed)
    wds = telemetry.WarpDeltaStatistics3d(float(rng.random()), float(rng.random()), float(rng.random()), float(rng.random()),
                                          float(rng.random()), telemetry.Vector3i(1 + seed, 2, 3), False, seed % 2 == 0)
    tds = telemetry.TsdfDifferenceStatistics3d(float(rng.random()), float(rng.random()), float(rng.random()),
                                               float(rng.random()), telemetry.Vector3i(4, 5 + seed, 6))
    return telemetry.ConvergenceReport3d(iterations, limit_reached, wds, tds)


def test_pair_cache_names_order_and_range(tmp_path):
    data = tmp_path / "data"
    canonical = np.arange(8, dtype=np.float32).reshape(2, 2, 2)
    for frame, row in ((2, 214), (10, 214), (2, 30)):
        path = multipair.save_pair(str(data), frame, row, canonical * frame, canonical + row)
        assert os.path.basename(path) == "data_%d_%d.npz" % (frame, row)
    os.makedirs(data / "images")
    entries = multipair.list_pair_cache(str(data))
    # the reference sorts the FILE NAMES (lexicographic): data_10_214 < data_2_214 < data_2_30
    assert [(f, r) for f, r, _ in entries] == [(10, 214), (2, 214), (2, 30)]
    assert [(f, r) for f, r, _ in multipair.list_pair_cache(str(data), 1, 2)] == [(2, 214)]
    loaded_canonical, loaded_live = multipair.load_pair(entries[0][2])
    assert loaded_canonical.dtype == np.float32 and loaded_canonical.flags["C_CONTIGUOUS"]
    assert np.array_equal(loaded_canonical, canonical * 10) and np.array_equal(loaded_live, canonical + 214)
    assert multipair.infer_frame_number_and_pixel_row_from_filename("data_17_233.npz") == (17, 233)


def test_report_table_columns_and_analysis(tmp_path):
    pytest.importorskip("pandas")
    report_sets = [[make_report(100, True, 0), make_report(7, False, 1)], [make_report(3, False, 2), make_report(100, True, 3)],
                   [make_report(5, False, 4), make_report(9, False, 5)]]
    pairs = [(10, 214), (11, 214), (12, 300)]
    frame = multipair.post_process_convergence_report_sets(report_sets, pairs)
    expected_columns = ["canonical_frame", "pixel_row"] + ["l%d_%s" % (level, c) for level in range(2)
                                                            for c in REFERENCE_LEVEL_COLUMNS]
    assert list(frame.columns) == expected_columns
    assert multipair.infer_level_count(frame) == 2
    assert list(frame["canonical_frame"]) == [10, 11, 12] and list(frame["pixel_row"]) == [214, 214, 300]
    assert list(frame["l0_iter_count"]) == [100, 3, 5] and list(frame["l1_iter_lim_reached"]) == [False, True, False]
    assert frame["l0_warp_delta_max_x"][1] == report_sets[1][0].warp_delta_statistics.longest_warp_location.x
    assert frame["l1_diff_max_y"][2] == report_sets[2][1].tsdf_difference_statistics.biggest_difference_location.y
    assert multipair.get_converged_ratio_for_level(frame, 0) == pytest.approx(2 / 3)
    assert multipair.get_mean_iteration_count_for_level(frame, 1) == pytest.approx((7 + 100 + 9) / 3)
    written = multipair.write_reports(frame, str(tmp_path))
    assert os.path.basename(written[0]) == "convergence_reports.pk" and all(os.path.exists(p) for p in written)
    import pandas as pd
    assert pd.read_pickle(written[0]).equals(frame)
    log = multipair.analyze_convergence_data(frame, str(tmp_path))
    text = open(log).read()
    # the reference's layout (run_hierarchical_optimizer3d_multipair.py:170-183): all levels of a statistic on one line
    assert text == ("Per-level convergence ratios:\n  level 0: 66.67%  level 1: 66.67%\n"
                    "Per-level mean iteration counts:\n  level 0: 36.00  level 1: 38.67\n")


def test_telemetry_log_file_layout_and_round_trip(tmp_path):
    """`telemetry_log.npz` in the reference's layout (hierarchical_optimization_visualizer.py:210-250): per level the keys
    l<i>_warp_fields / l<i>_data_term_gradients / l<i>_tikhonov_term_gradients holding the iterations' [H][W][2] fields
    stacked along axis 2, an empty array where a kind was not collected; the folder name of a pair
    (run_hierarchical_optimizer3d_multipair.py:214-215)"""
    from lsf_b200.hierarchical import OptimizationIterationData2d
    rng = np.random.default_rng(5)
    log = []
    for size, frames in ((4, 3), (8, 2)):
        level = OptimizationIterationData2d()
        for _ in range(frames):
            level.add_iteration_result(rng.random((size, size), dtype=np.float32), rng.random((size, size, 2), dtype=np.float32),
                                       rng.random((size, size, 2), dtype=np.float32), None)
        log.append(level)
    metadata = multipair.get_telemetry_metadata(log)
    assert (metadata.has_warp_fields, metadata.has_data_term_gradients, metadata.has_tikhonov_term_gradients,
            metadata.field_size) == (True, True, False, 4)
    folder = multipair.get_telemetry_subfolder_path(str(tmp_path / "telemetry"), 12, 214)
    assert os.path.basename(folder) == "pair_12-13_214"
    multipair.save_telemetry_log(log, metadata, folder)
    stored = np.load(os.path.join(folder, "telemetry_log.npz"))
    assert sorted(stored.files) == sorted("l%d_%s" % (i, kind) for i in (0, 1)
                                          for kind in ("warp_fields", "data_term_gradients", "tikhonov_term_gradients"))
    assert stored["l0_warp_fields"].shape == (4, 4, 6) and stored["l1_warp_fields"].shape == (8, 8, 4)
    assert stored["l0_tikhonov_term_gradients"].size == 0
    assert np.array_equal(stored["l0_warp_fields"][:, :, 2:4], log[0].get_warp_fields()[1])
    loaded = multipair.load_telemetry_log(folder)
    assert [level.get_frame_count() for level in loaded] == [3, 2]
    for level, original in zip(loaded, log):
        for a, b in zip(level.get_warp_fields(), original.get_warp_fields()):
            assert np.array_equal(a, b)
        for a, b in zip(level.get_data_term_gradients(), original.get_data_term_gradients()):
            assert np.array_equal(a, b)
        assert all(t is None for t in level.get_tikhonov_term_gradients())
