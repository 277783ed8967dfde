#!/usr/bin/env python
"""Generates the golden fixtures under tests/golden/ from the reference checkout.

Run in the build container only (needs /root/reference; the GPU box never runs this):

    python tests/golden/make_golden.py

Produces
  reference_literals.npz   every literal matrix / tensor / vector found in the reference's own tests:
                           * tests/test_data/*.py (numpy literals, imported),
                           * cpp/tests/*.cpp and cpp/tests/data/*.hpp (Eigen comma-initialisers and
                             Tensor::setValues blocks, parsed textually).
                           Key format:  "<file stem>/<test case or static name>/<variable>[#k]"
  reference_slavcheva_runs.npz  per-voxel Killing / Tikhonov / level-set / data terms and whole runs of the reference's
                           Python SlavchevaOptimizer2d (DIRECT), see collect_slavcheva_runs().
  reference_python_runs.npz  inputs + outputs obtained by RUNNING the reference's Python implementation
                           (math_utils, nonrigid_opt) on seeded inputs, with the C++ extension / matplotlib
                           imports stubbed out (the reference Python is the only runnable reference here,
                           SURVEY.md section 8c).

Only data (numbers) is extracted; no reference source code is copied.
"""
import os
import re
import sys
import types

import numpy as np

REF = os.environ.get("LSF_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))

NUMBER = re.compile(r"(?<![A-Za-z_0-9.])[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?[fF]?(?![A-Za-z_0-9.])")
IDENT = re.compile(r"(?<![0-9.A-Za-z_])[A-Za-z_][A-Za-z_0-9:]*")


def strip_comments(text):
    text = re.sub(r"/\*.*?\*/", lambda m: " " * len(m.group(0)), text, flags=re.S)
    return re.sub(r"//[^\n]*", lambda m: " " * len(m.group(0)), text)


def numbers_in(chunk):
    chunk = IDENT.sub(" ", chunk)
    return [float(t.rstrip("fF")) for t in NUMBER.findall(chunk)]


def trailing_dims(type_text):
    t = type_text.replace(" ", "")
    if "m2f" in t or "Matrix2f" in t:
        return (2, 2)
    if "m3f" in t or "Matrix3f" in t:
        return (3, 3)
    if "v2f" in t or "Vector2f" in t:
        return (2,)
    if "v3f" in t or "Vector3f" in t:
        return (3,)
    return ()


def parse_cpp_literals(path):
    """Returns {key: ndarray} for every `name << ...;` / `name.setValues(...);` whose element count matches
    the shape the variable was declared with."""
    raw = open(path).read()
    text = strip_comments(raw)
    stem = os.path.splitext(os.path.basename(path))[0]
    scopes = [(m.start(), m.group(1)) for m in re.finditer(r"BOOST_AUTO_TEST_CASE\s*\(\s*(\w+)\s*\)", text)]
    scopes += [(m.start(), m.group(1)) for m in re.finditer(r"\bstatic\s+[\w:<>, ]+?\s+(\w+)\s*=", text)]
    scopes.sort()

    def scope_at(pos):
        name = "global"
        for start, scope_name in scopes:
            if start <= pos:
                name = scope_name
            else:
                break
        return name

    out = {}
    counts = {}
    for m in re.finditer(r"\b(\w+)\s*(<<|\.setValues\s*\()", text):
        name = m.group(1)
        if name in ("cout", "cerr", "std", "ss"):
            continue
        # data chunk: up to the terminating ';'
        end = text.find(";", m.end())
        chunk = text[m.end():end]
        # declaration: nearest preceding "name(d0, d1, ...)"
        decl = None
        for d in re.finditer(r"\b" + re.escape(name) + r"\s*\(\s*(\d+(?:\s*,\s*\d+)*)\s*\)", text[:m.start()]):
            decl = d
        if decl is None:
            continue
        dims = tuple(int(v) for v in decl.group(1).split(","))
        stmt_start = max(text.rfind(";", 0, decl.start()), text.rfind("{", 0, decl.start()),
                         text.rfind("}", 0, decl.start())) + 1
        type_text = text[stmt_start:decl.start()]
        shape = dims + trailing_dims(type_text)
        values = numbers_in(chunk)
        if len(values) != int(np.prod(shape)):
            continue
        key = "%s/%s/%s" % (stem, scope_at(m.start()), name)
        k = counts.get(key, 0)
        counts[key] = k + 1
        if k:
            key = "%s#%d" % (key, k)
        out[key] = np.array(values, dtype=np.float32).reshape(shape)
    # `float data[] = {...}` blocks (e.g. pyramid3d_argument_field)
    for m in re.finditer(r"float\s+(\w+)\s*\[\s*\]\s*=\s*\{(.*?)\}\s*;", text, flags=re.S):
        key = "%s/%s/%s" % (stem, scope_at(m.start()), m.group(1))
        out[key] = np.array(numbers_in(m.group(2)), dtype=np.float32)
    # scalar statics: `static float name = value;`
    for m in re.finditer(r"static\s+float\s+(\w+)\s*=\s*([-+0-9.eE]+)f?\s*;", text):
        out["%s/%s/value" % (stem, m.group(1))] = np.array(float(m.group(2)), dtype=np.float32)
    return out


def collect_literals():
    out = {}
    sys.path.insert(0, REF)
    import tests.test_data.hierarchical_optimizer_test_data as hier_data
    import tests.test_data.test_data_convolution as conv_data
    for module, prefix in ((hier_data, "py_hierarchical"), (conv_data, "py_convolution")):
        for name in dir(module):
            value = getattr(module, name)
            if isinstance(value, np.ndarray):
                out["%s/%s" % (prefix, name)] = value
    cpp_tests = os.path.join(REF, "cpp", "tests")
    for fname in sorted(os.listdir(cpp_tests)):
        if fname.endswith(".cpp"):
            out.update(parse_cpp_literals(os.path.join(cpp_tests, fname)))
    for fname in sorted(os.listdir(os.path.join(cpp_tests, "data"))):
        if fname.endswith(".hpp"):
            out.update(parse_cpp_literals(os.path.join(cpp_tests, "data", fname)))
    return out


# ----------------------------------------------------------------------------------------------------------
def install_stubs():
    """Stub modules so the reference's Python optimizers import without the C++ extension / matplotlib."""
    class _Anything(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            sub = _Anything(self.__name__ + "." + name)
            setattr(self, name, sub)
            return sub

        def __call__(self, *args, **kwargs):
            return _Anything("call")

    for name in ("level_set_fusion_optimization", "matplotlib", "matplotlib.pyplot", "matplotlib.cm",
                 "matplotlib.colors", "mpl_toolkits", "mpl_toolkits.axes_grid1", "sktensor", "progressbar",
                 "matplotlib.patches", "matplotlib.gridspec", "matplotlib.ticker"):
        if name not in sys.modules:
            sys.modules[name] = _Anything(name)


def synthetic_pair_2d(size, shift=2.0):
    """Deterministic 2D TSDF pair: a circle and a line, live = shifted copy; half-width 10 voxels."""
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)

    def sdf(cx, cy, r, line):
        circle = np.sqrt((xx - cx) ** 2 + (yy - cy) ** 2) - r
        plane = line - yy
        return np.clip(np.minimum(circle, plane) / 10.0, -1.0, 1.0).astype(np.float32)

    canonical = sdf(size * 0.5, size * 0.45, size * 0.22, size * 0.8)
    live = sdf(size * 0.5 + shift, size * 0.45 - 0.6 * shift, size * 0.23, size * 0.8 - 1.3)
    return canonical, live


def collect_python_runs():
    install_stubs()
    sys.path.insert(0, REF)
    out = {}
    rng = np.random.default_rng(20181218)

    import math_utils.convolution as mc
    kernel7 = mc.sobolev_kernel_1d.astype(np.float32)
    out["kernel7"] = kernel7
    v2d = rng.standard_normal((12, 12, 2)).astype(np.float32)
    out["conv2d/in"] = v2d.copy()
    res = v2d.copy()
    mc.convolve_with_kernel(res, kernel7)
    out["conv2d/out"] = res
    v2dz = v2d.copy()
    v2dz[rng.random((12, 12)) < 0.4] = 0.0
    out["conv2d_preserve_zeros/in"] = v2dz.copy()
    res = v2dz.copy()
    mc.convolve_with_kernel_preserve_zeros(res, kernel7)
    out["conv2d_preserve_zeros/out"] = res
    v3d = rng.standard_normal((9, 10, 11, 3)).astype(np.float32)
    out["conv3d/in"] = v3d.copy()
    res = v3d.copy()
    mc.convolve_with_kernel(res, kernel7)
    out["conv3d/out"] = res

    import math_utils.resampling as mr
    s3d = rng.standard_normal((6, 8, 10)).astype(np.float32)
    out["resample3d/in"] = s3d
    out["resample3d/up_linear"] = mr.upsample2x_linear(s3d).astype(np.float32)
    out["resample3d/down_linear"] = mr.downsample2x_linear(s3d).astype(np.float32)

    import nonrigid_opt.field_warping as fw
    f2d = rng.random((16, 16)).astype(np.float32) * 2 - 1
    w2d = (rng.standard_normal((16, 16, 2)) * 1.5).astype(np.float32)
    out["warp2d/field"] = f2d
    out["warp2d/warp"] = w2d
    out["warp2d/out"] = fw.warp_field(f2d, w2d).astype(np.float32)
    out["warp2d/out_replacement0"] = fw.warp_field_replacement(f2d, w2d, 0.0).astype(np.float32)

    from nonrigid_opt.hierarchical.pyramid import ScalarFieldPyramid2d
    pyr = ScalarFieldPyramid2d(f2d, 8)
    for i, level in enumerate(pyr.levels):
        out["pyramid2d/level%d" % i] = np.asarray(level, dtype=np.float32)

    # the reference's Python hierarchical optimizer with Tikhonov term + Sobolev kernel enabled
    # (unpinned by the reference's own fixtures): 32x32, chunk 4 -> levels 8,16,32
    import nonrigid_opt.hierarchical.hierarchical_optimizer2d as ho
    canonical, live = synthetic_pair_2d(32)
    out["hier2d_full/canonical"] = canonical
    out["hier2d_full/live"] = live
    for tag, kwargs in (
            ("data_only", dict(tikhonov_term_enabled=False, gradient_kernel_enabled=False)),
            # strength 0.1: the Laplacian feedback (F4) amplifies the 2D checkerboard mode by 8*strength per
            # iteration, so 0.2 without a kernel diverges within a few iterations
            ("tikhonov", dict(tikhonov_term_enabled=True, gradient_kernel_enabled=False, tikhonov_strength=0.1)),
            ("tikhonov_kernel", dict(tikhonov_term_enabled=True, gradient_kernel_enabled=True,
                                     tikhonov_strength=0.2, kernel=kernel7))):
        optimizer = ho.HierarchicalOptimizer2d(maximum_chunk_size=4, rate=0.1, maximum_iteration_count=25,
                                               maximum_warp_update_threshold=0.001, data_term_amplifier=1.0,
                                               **kwargs)
        warp = optimizer.optimize(canonical.copy(), live.copy())
        out["hier2d_full/%s/warp" % tag] = np.asarray(warp, dtype=np.float32)
    return out


def collect_slavcheva_runs():
    """Runs of the reference's Python SlavchevaOptimizer2d (ComputeMethod.DIRECT: the only implementation of the
    Killing and level-set terms) and of its per-voxel term functions on seeded inputs. The visualizer is replaced by
    a no-op and convergence-status logging (a C++-extension call) is off; the optimizer writes nothing else."""
    install_stubs()
    sys.path.insert(0, REF)
    import contextlib
    import io
    import tempfile
    out = {}
    rng = np.random.default_rng(20190115)
    import utils.sampling as sampling
    sampling.set_focus_coordinates(-5, -5)  # no voxel matches: suppress the per-voxel debug prints
    import nonrigid_opt.slavcheva.smoothing_term as st
    import nonrigid_opt.slavcheva.data_term as dt
    import nonrigid_opt.slavcheva.level_set_term as ls
    import nonrigid_opt.slavcheva.slavcheva_optimizer2d as so
    import nonrigid_opt.slavcheva.slavcheva_visualizer as viz

    class NoVisualizer:
        Settings = viz.SlavchevaVisualizer.Settings

        def __init__(self, *args, **kwargs):
            self.data_component_field = None
            self.smoothing_component_field = None
            self.level_set_component_field = None

        def write_live_sdf_visualizations(self, *args, **kwargs):
            pass

        def write_all_iteration_visualizations(self, *args, **kwargs):
            pass

    so.viz.SlavchevaVisualizer = NoVisualizer

    # ---- per-voxel terms on random fields
    size = 9
    warp = (rng.standard_normal((size, size, 2)) * 0.3).astype(np.float32)
    live = np.clip(rng.standard_normal((size, size)) * 0.6, -1, 1).astype(np.float32)
    out["terms/warp"] = warp
    out["terms/live"] = live
    killing = np.zeros_like(warp)
    tikhonov = np.zeros_like(warp)
    level_set = np.zeros_like(warp)
    for y in range(size):
        for x in range(size):
            killing[y, x] = st.compute_local_smoothing_term_gradient_killing(warp, x, y, copy_if_zero=False,
                                                                             isomorphic_enforcement_factor=0.1)[0]
            tikhonov[y, x] = st.compute_local_smoothing_term_gradient_tikhonov(warp, x, y, copy_if_zero=False)[0]
            level_set[y, x] = ls.level_set_term_at_location(live, x, y)[0]
    out["terms/killing_lambda0.1"] = killing
    out["terms/tikhonov_direct"] = tikhonov
    out["terms/level_set"] = level_set
    canonical = np.clip(live + (rng.standard_normal((size, size)) * 0.2).astype(np.float32), -1, 1).astype(np.float32)
    out["terms/canonical"] = canonical
    gy, gx = np.gradient(live)
    basic = np.zeros_like(warp)
    fdm = np.zeros_like(warp)
    for y in range(size):
        for x in range(size):
            basic[y, x] = dt.compute_local_data_term_gradient_basic(live, canonical, x, y, gx, gy)[0]
            fdm[y, x] = dt.compute_local_data_term_gradient_thresholded_fdm(live, canonical, x, y, gx, gy)[0]
    out["terms/data_basic"] = basic
    out["terms/data_thresholded_fdm"] = fdm
    # energy aggregates of ComputeMethod.VECTORIZED (slavcheva_optimizer2d.py:169-175; the run itself needs the C++ extension)
    # on the same random fields: [data energy, smoothing energy] before the weights
    out["terms/vectorized_energies"] = np.array([dt.compute_data_term_energy_contribution(live.copy(), canonical),
                                                 st.compute_smoothing_term_energy(warp, live, canonical)], dtype=np.float64)

    # ---- optimizer runs (32x32 synthetic pair, 7-tap Sobolev kernel of the reference)
    import math_utils.convolution as mc
    kernel7 = mc.sobolev_kernel_1d.astype(np.float32)
    canonical, live = synthetic_pair_2d(32)
    out["runs/canonical"] = canonical
    out["runs/live"] = live
    out["runs/kernel7"] = kernel7
    cases = {
        "tikhonov_sobolev": dict(smoothing_term_method=st.SmoothingTermMethod.TIKHONOV, level_set_term_enabled=False,
                                 sobolev_smoothing_enabled=True),
        "killing_levelset_sobolev": dict(smoothing_term_method=st.SmoothingTermMethod.KILLING,
                                         level_set_term_enabled=True, sobolev_smoothing_enabled=True),
        "killing_levelset_plain": dict(smoothing_term_method=st.SmoothingTermMethod.KILLING,
                                       level_set_term_enabled=True, sobolev_smoothing_enabled=False),
        "fdm_tikhonov_sobolev": dict(smoothing_term_method=st.SmoothingTermMethod.TIKHONOV,
                                     data_term_method=dt.DataTermMethod.THRESHOLDED_FDM, level_set_term_enabled=False,
                                     sobolev_smoothing_enabled=True),
    }
    with tempfile.TemporaryDirectory() as scratch:
        for tag, kwargs in cases.items():
            for iterations in (1, 5):
                optimizer = so.SlavchevaOptimizer2d(out_path=scratch, field_size=32,
                                                    compute_method=so.ComputeMethod.DIRECT,
                                                    gradient_descent_rate=0.1, data_term_weight=1.0,
                                                    smoothing_term_weight=0.2, isomorphic_enforcement_factor=0.1,
                                                    level_set_term_weight=0.2,
                                                    maximum_warp_length_lower_threshold=0.001,
                                                    maximum_warp_length_upper_threshold=10000,
                                                    max_iterations=iterations, min_iterations=1,
                                                    sobolev_kernel=kernel7, enable_convergence_status_logging=False,
                                                    **kwargs)
                field = live.copy()
                with contextlib.redirect_stdout(io.StringIO()):
                    optimizer.optimize(field, canonical.copy())
                out["runs/%s/live_after_%d" % (tag, iterations)] = field.astype(np.float32)
                out["runs/%s/max_warps_%d" % (tag, iterations)] = np.array(optimizer.log.max_warps, dtype=np.float32)
                # OptimizationLog energies (slavcheva_optimizer2d.py:370-374): one row per iteration
                out["runs/%s/energies_%d" % (tag, iterations)] = np.array(
                    [optimizer.log.data_energies, optimizer.log.smoothing_energies, optimizer.log.level_set_energies],
                    dtype=np.float64).T
    return out


def main():
    if "--slavcheva-only" in sys.argv:
        runs = collect_slavcheva_runs()
        np.savez_compressed(os.path.join(OUT, "reference_slavcheva_runs.npz"), **runs)
        print("reference_slavcheva_runs.npz: %d arrays" % len(runs))
        return
    literals = collect_literals()
    np.savez_compressed(os.path.join(OUT, "reference_literals.npz"), **literals)
    print("reference_literals.npz: %d arrays" % len(literals))
    runs = collect_python_runs()
    np.savez_compressed(os.path.join(OUT, "reference_python_runs.npz"), **runs)
    print("reference_python_runs.npz: %d arrays" % len(runs))
    runs = collect_slavcheva_runs()
    np.savez_compressed(os.path.join(OUT, "reference_slavcheva_runs.npz"), **runs)
    print("reference_slavcheva_runs.npz: %d arrays" % len(runs))


if __name__ == "__main__":
    main()
