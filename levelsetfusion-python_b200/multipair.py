"""The reference's multi-pair experiment driver and its on-disk formats, on top of
multigpu.optimize_pairs: reference run_hierarchical_optimizer3d_multipair.py.

* pair cache     -- one `data_<canonical frame>_<pixel row>.npz` per pair with arrays `canonical` and `live`
                    (reference :320-324 writes them, :341-355 reads them back: os.listdir + sort, an `images` entry is
                    skipped, frame number and pixel row are the first two digit groups of the file name, :73-83)
* report table   -- a pandas DataFrame with `canonical_frame`, `pixel_row` and 17 columns per pyramid level
                    (reference post_process_convergence_report_sets, :85-131), pickled as `convergence_reports.pk`
                    (:437-441; the reference also writes .xlsx, which needs openpyxl -- written when it is importable,
                    a .csv otherwise)
* analysis       -- per-level converged ratio and mean iteration count (reference :135-175 -> `analysis.txt`)

The pairs are sharded over the ranks of torch.distributed (none: one rank) and, inside a rank, over `streams` worker
threads; the convergence reports are gathered on every rank, rank 0 writes the files."""
import os
import re

import numpy as np

from . import multigpu

_DIGITS = re.compile(r"\d+")
LEVEL_COLUMNS = ("iter_count", "iter_lim_reached", "warp_delta_amt_ratio", "warp_delta_min", "warp_delta_max",
                 "warp_delta_mean", "warp_delta_std", "warp_delta_max_x", "warp_delta_max_y", "warps_below_min_thresh",
                 "warps_above_max_thresh", "diff_delta_min", "diff_delta_max", "diff_delta_mean", "diff_delta_std",
                 "diff_max_x", "diff_max_y")


def infer_frame_number_and_pixel_row_from_filename(filename):
    """reference :80-83"""
    found = _DIGITS.findall(filename)
    return int(found[0]), int(found[1])


def pair_file_name(canonical_frame, pixel_row):
    """reference :323 (np.savez appends .npz)"""
    return "data_{:d}_{:d}.npz".format(int(canonical_frame), int(pixel_row))


def save_pair(data_path, canonical_frame, pixel_row, canonical_field, live_field):
    os.makedirs(data_path, exist_ok=True)
    path = os.path.join(data_path, pair_file_name(canonical_frame, pixel_row))
    np.savez(path, canonical=np.asarray(canonical_field, dtype=np.float32), live=np.asarray(live_field, dtype=np.float32))
    return path


def list_pair_cache(data_path, start_from_index=0, stop_before_index=10000000):
    """[(canonical frame, pixel row, path)] in the reference's order (sorted file names, `images` skipped, then the
    start/stop range, reference :341-360)"""
    files = sorted(f for f in os.listdir(data_path) if f != "images")
    entries = []
    for name in files:
        frame, row = infer_frame_number_and_pixel_row_from_filename(name)
        entries.append((frame, row, os.path.join(data_path, name)))
    return entries[start_from_index:min(len(entries), stop_before_index)]


def load_pair(path):
    """-> (canonical, live), float32 C-contiguous (what the optimizers accept)"""
    with np.load(path) as archive:
        return (np.ascontiguousarray(archive["canonical"], dtype=np.float32),
                np.ascontiguousarray(archive["live"], dtype=np.float32))


def post_process_convergence_report_sets(convergence_report_sets, frame_numbers_and_rows):
    """reference :85-131: same column names, same order (the two pair columns come first, then the levels)"""
    import pandas as pd
    data = {"canonical_frame": [], "pixel_row": []}
    level_count = len(convergence_report_sets[0]) if convergence_report_sets else 0
    for level in range(level_count):
        for column in LEVEL_COLUMNS:
            data["l%d_%s" % (level, column)] = []
    for report_set, (frame_number, pixel_row) in zip(convergence_report_sets, frame_numbers_and_rows):
        data["canonical_frame"].append(frame_number)
        data["pixel_row"].append(pixel_row)
        for level, report in enumerate(report_set):
            wds, tds = report.warp_delta_statistics, report.tsdf_difference_statistics
            values = (report.iteration_count, report.iteration_limit_reached, wds.ratio_above_min_threshold,
                      wds.length_min, wds.length_max, wds.length_mean, wds.length_standard_deviation,
                      wds.longest_warp_location.x, wds.longest_warp_location.y, wds.is_largest_below_min_threshold,
                      wds.is_largest_above_max_threshold, tds.difference_min, tds.difference_max, tds.difference_mean,
                      tds.difference_standard_deviation, tds.biggest_difference_location.x,
                      tds.biggest_difference_location.y)
            for column, value in zip(LEVEL_COLUMNS, values):
                data["l%d_%s" % (level, column)].append(value)
    return pd.DataFrame.from_dict(data)


def infer_level_count(data_frame):
    """reference :154-161"""
    return (len(data_frame.columns) - 2) // len(LEVEL_COLUMNS)


def get_converged_ratio_for_level(data_frame, level):
    """reference :135-143"""
    reached = data_frame["l%d_iter_lim_reached" % level].astype(bool)
    return 0.0 if len(data_frame) == 0 else float((~reached).sum()) / len(data_frame)


def get_mean_iteration_count_for_level(data_frame, level):
    """reference :146-148"""
    return float(data_frame["l%d_iter_count" % level].mean())


def analyze_convergence_data(data_frame, out_path):
    """reference :164-185: analysis.txt with the per-level convergence ratios and mean iteration counts; the file has the
    reference's layout byte for byte (all levels of a statistic on ONE line, each as "  level <i>: <value>")"""
    log_path = os.path.join(out_path, "analysis.txt")
    levels = range(infer_level_count(data_frame))
    lines = ["Per-level convergence ratios:",
             "".join("  level {:d}: {:.2%}".format(level, get_converged_ratio_for_level(data_frame, level)) for level in levels),
             "Per-level mean iteration counts:",
             "".join("  level {:d}: {:.2f}".format(level, get_mean_iteration_count_for_level(data_frame, level))
                     for level in levels)]
    with open(log_path, "w") as log_file:
        log_file.write("\n".join(lines) + "\n")
    return log_path


def write_reports(data_frame, out_path, reports_file_name="convergence_reports"):
    """reference :437-441"""
    os.makedirs(out_path, exist_ok=True)
    written = [os.path.join(out_path, reports_file_name + ".pk")]
    data_frame.to_pickle(written[0])
    try:
        import openpyxl  # noqa: F401
        data_frame.to_excel(os.path.join(out_path, reports_file_name + ".xlsx"))
        written.append(os.path.join(out_path, reports_file_name + ".xlsx"))
    except ImportError:
        data_frame.to_csv(os.path.join(out_path, reports_file_name + ".csv"))
        written.append(os.path.join(out_path, reports_file_name + ".csv"))
    return written


# ------------------------------------------------------------------------------------------------ telemetry_log.npz
def get_telemetry_subfolder_path(telemetry_folder, frame_number, pixel_row):
    """reference run_hierarchical_optimizer3d_multipair.py:214-215"""
    return os.path.join(telemetry_folder, "pair_{:d}-{:d}_{:d}".format(frame_number, frame_number + 1, pixel_row))


class TelemetryMetadata:
    """reference hierarchical_optimization_visualizer.py:150-157"""

    def __init__(self, has_warp_fields, has_data_term_gradients, has_tikhonov_term_gradients, field_size):
        self.has_warp_fields = has_warp_fields
        self.has_data_term_gradients = has_data_term_gradients
        self.has_tikhonov_term_gradients = has_tikhonov_term_gradients
        self.field_size = field_size


def get_telemetry_metadata(telemetry_log):
    """what the first level of a per-level iteration-data list holds (reference
    hierarchical_optimization_visualizer.py:182-208; that code files a non-empty Tikhonov list under
    has_data_term_gradients -- here it sets has_tikhonov_term_gradients)"""
    first = telemetry_log[0]

    def probe(fields):
        if fields is not None and len(fields) > 0 and fields[0] is not None and np.asarray(fields[0]).size > 0:
            return True, np.asarray(fields[0]).shape[0]
        return False, 0

    has_warp, size_warp = probe(first.get_warp_fields())
    has_data, size_data = probe(first.get_data_term_gradients())
    has_tikhonov, size_tikhonov = probe(first.get_tikhonov_term_gradients())
    return TelemetryMetadata(has_warp, has_data, has_tikhonov, size_warp or size_data or size_tikhonov)


def save_telemetry_log(telemetry_log, telemetry_metadata, output_folder):
    """`telemetry_log.npz` of one pair in the reference's layout (hierarchical_optimization_visualizer.py:210-227): per
    level i the keys l<i>_warp_fields, l<i>_data_term_gradients, l<i>_tikhonov_term_gradients, each the fields of all
    iterations stacked along axis 2 (np.dstack) or an empty array. `telemetry_log` = get_per_level_iteration_data()."""
    telemetry_dict = {}
    for i_level, level_data in enumerate(telemetry_log):
        telemetry_dict["l{:d}_warp_fields".format(i_level)] = np.dstack(level_data.get_warp_fields())
        telemetry_dict["l{:d}_data_term_gradients".format(i_level)] = \
            np.array([]) if not telemetry_metadata.has_data_term_gradients else np.dstack(
                level_data.get_data_term_gradients())
        telemetry_dict["l{:d}_tikhonov_term_gradients".format(i_level)] = \
            np.array([]) if not telemetry_metadata.has_tikhonov_term_gradients else np.dstack(
                level_data.get_tikhonov_term_gradients())
    os.makedirs(output_folder, exist_ok=True)
    np.savez_compressed(os.path.join(output_folder, "telemetry_log.npz"), **telemetry_dict)


def load_telemetry_log(output_folder, components=2):
    """inverse of save_telemetry_log (reference :230-250, which splits the stacks into 2-component fields: written for 2D
    runs -- 3D fields [X][Y][Z][3] are stacked along their Z axis by np.dstack and split back with components = Z). Returns
    a list of OptimizationIterationData2d (the live fields are not part of the file)."""
    from .hierarchical import OptimizationIterationData2d
    telemetry_dict = np.load(os.path.join(output_folder, "telemetry_log.npz"))
    level_count = len(telemetry_dict.files) // 3
    telemetry_log = []
    for i_level in range(level_count):
        def split(key):
            stack = telemetry_dict["l{:d}_{:s}".format(i_level, key)]
            if stack.ndim < 3 or stack.size == 0:
                return []
            return [np.ascontiguousarray(part) for part in np.dsplit(stack, stack.shape[2] // components)]
        warp_fields, data_term_gradients = split("warp_fields"), split("data_term_gradients")
        tikhonov_term_gradients = split("tikhonov_term_gradients")
        level_data = OptimizationIterationData2d()
        for i, warp_field in enumerate(warp_fields):
            level_data.add_iteration_result(None, warp_field, data_term_gradients[i] if data_term_gradients else None,
                                            tikhonov_term_gradients[i] if tikhonov_term_gradients else None)
        telemetry_log.append(level_data)
    return telemetry_log


def run_multipair(data_path, out_path, optimizer_factory, streams=1, start_from_index=0, stop_before_index=10000000,
                  save_warps=False, group=None, save_telemetry=False):
    """Optimises every cached pair (reference :403-436) over the ranks / streams and writes the report table.

    optimizer_factory -- callable() -> HierarchicalOptimizer3d with
                         LoggingParameters(collect_per_level_convergence_reports=True) (one object per worker thread)
    save_warps        -- also write `warp_<frame>_<row>.npy` next to the reports (this rank's pairs)
    save_telemetry    -- reference --save_telemetry (:395-411,443-453): the optimizers must be built with
                         LoggingParameters(collect_per_level_iteration_data=True); every pair's per-level iteration data
                         goes to <out_path>/telemetry/pair_<f>-<f+1>_<row>/telemetry_log.npz (written by the pair's rank)
    Returns the DataFrame (all ranks)."""
    entries = list_pair_cache(data_path, start_from_index, stop_before_index)
    rank, world_size, _ = multigpu.world()
    if save_warps:
        os.makedirs(out_path, exist_ok=True)

    def call(optimizer, canonical, live, entry=None):
        warp = optimizer.optimize(canonical, live)
        telemetry_log = optimizer.get_per_level_iteration_data() if save_telemetry else None
        return warp, optimizer.get_per_level_convergence_reports(), telemetry_log

    def one_pair(optimizer, canonical, live):
        return call(optimizer, canonical, live)

    worker = multigpu.PerWorkerOptimizer(optimizer_factory, one_pair)
    local = {}

    def load(index):
        return load_pair(entries[index][2])

    results = multigpu.optimize_pairs(worker, len(entries), load, rank, world_size, gather=False, streams=streams)
    for index, (warp, reports, telemetry_log) in results.items():
        frame, row, _ = entries[index]
        if save_warps:
            np.save(os.path.join(out_path, "warp_{:d}_{:d}.npy".format(frame, row)), np.asarray(warp))
        if save_telemetry and telemetry_log:
            save_telemetry_log(telemetry_log, get_telemetry_metadata(telemetry_log),
                               get_telemetry_subfolder_path(os.path.join(out_path, "telemetry"), frame, row))
        local[index] = reports
    if world_size > 1:
        import torch.distributed as dist
        shards = [None] * world_size
        dist.all_gather_object(shards, local, group=group)
        merged = {}
        for shard in shards:
            merged.update(shard)
    else:
        merged = local
    report_sets = [merged[i] for i in range(len(entries))]
    frame_numbers_and_rows = [(frame, row) for frame, row, _ in entries]
    data_frame = post_process_convergence_report_sets(report_sets, frame_numbers_and_rows)
    if rank == 0:
        write_reports(data_frame, out_path)
        analyze_convergence_data(data_frame, out_path)
    return data_frame
