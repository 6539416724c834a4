"""CPU-only tests: host-side logic against the reference goldens, the C-ABI surface, multi-rank merge (gloo)."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle as orc


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads without a GPU and exports exactly what include/pypairs_b200.h declares."""
    import __graft_entry__ as g
    g.build()
    from pypairs_b200 import _ffi
    lib = _ffi.load()
    header = open(os.path.join(ROOT, "include", "pypairs_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(pp_[a-z_0-9]+)\s*\(", header)))
    assert declared == sorted(_ffi.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.pp_version() == int(re.search(r"#define PP_ABI_VERSION (\d+)", header).group(1)) == 2
    h = ctypes.c_void_p()
    if lib.pp_device_count() == 0:  # no compute without a GPU: creation fails loudly, no fallback
        assert lib.pp_ctx_create(0, ctypes.byref(h)) != 0
        assert len(lib.pp_last_error(None)) > 0
        with pytest.raises(_ffi.PyPairsB200Error):
            _ffi.sandbag(np.ones((4, 4)), [0, 1], [1, 1], np.arange(4), [1, 1])


def test_ctypes_prototypes_match_the_header():
    """Every prototype of include/pypairs_b200.h has as many parameters as the ctypes binding declares."""
    from pypairs_b200 import _ffi
    lib = _ffi.load()
    header = open(os.path.join(ROOT, "include", "pypairs_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    protos = re.findall(r"\b(pp_[a-z_0-9]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S)
    assert len(protos) >= len(_ffi.SYMBOLS)
    checked = 0
    for name, params in protos:
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        fn = getattr(lib, name)
        if fn.argtypes is not None:
            assert len(fn.argtypes) == n, (name, len(fn.argtypes), n)
            checked += 1
    assert checked >= 20


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pypairs_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert "libpp_oracle" not in src, f


def _emulated_sandbag(monkeypatch):
    """Route _ffi.sandbag to the CPU oracle so the HOST logic can be checked without a GPU (tests only)."""
    from pypairs_b200 import _ffi

    def fake(x, cell_rows, class_sizes, gene_cols, thresholds, shard=0, n_shards=1, **kw):
        if _ffi.is_sparse(x):
            assert x.format == "csr"  # the host keeps sparse input sparse and hands CSR to the library
            x = x.toarray()
        x = np.asarray(x, dtype=np.float64)
        gene_cols = np.asarray(gene_cols)
        if kw.get("drop_unexpressed"):
            gene_cols = gene_cols[(x != 0).any(axis=0)[gene_cols]]
        cats = np.concatenate([np.full(n, k) for k, n in enumerate(class_sizes)])
        res = orc.check_pairs(x[np.asarray(cell_rows)][:, gene_cols], cats, np.asarray(thresholds))
        r, c = np.where(res != -1)
        keys = (r * len(gene_cols) + c).astype(np.uint64)
        cls = res[r, c].astype(np.int32)
        sel = (np.arange(len(keys)) % n_shards) == shard
        return keys[sel], cls[sel], {}, gene_cols.astype(np.int32)

    monkeypatch.setattr(_ffi, "sandbag", fake)


def test_sandbag_host_logic_vs_reference(monkeypatch, sandbag_api, sandbag_minref):
    from pandas import DataFrame
    from pypairs_b200 import pairs
    _emulated_sandbag(monkeypatch)
    a = sandbag_api
    x = np.array(a["x"])
    sn, gn, lab = a["sample_names"], a["gene_names"], np.array(a["labels"])
    names3 = ["G1", "S", "G2M"]
    ann = {"int": {k: np.where(lab == i)[0].tolist() for i, k in enumerate(names3)},
           "str": {k: [sn[j] for j in np.where(lab == i)[0]] for i, k in enumerate(names3)},
           "bool": {k: (lab == i).tolist() for i, k in enumerate(names3)}}
    ann["int_rev"] = {k: ann["int"][k] for k in ["G2M", "G1", "S"]}
    ann["int_empty"] = dict(ann["int"], EMPTY=[58])
    for run in a["runs"]:
        kw = dict(run["kwargs"])
        annotation = ann[kw.pop("annotation")]
        if kw.pop("dataframe", False):
            got = pairs.sandbag(DataFrame(x, index=sn, columns=gn), annotation, **kw)
        else:
            got = pairs.sandbag(x, annotation, gn, sn, **kw)
        assert [[str(k), [[str(p), str(q)] for p, q in v]] for k, v in got.items()] == run["result"], run["tag"]
        assert all(type(k) is np.str_ for k in got.keys())  # keys and names are numpy strings (S8)
    m = sandbag_minref
    xm = np.array(m["matrix_genes_by_samples"]).T
    for frac_s, ref in m["results"].items():
        got = pairs.sandbag(xm, m["annotation"], m["gene_names"], m["sample_names"], fraction=float(frac_s))
        assert [[str(k), [[str(p), str(q)] for p, q in v]] for k, v in got.items()] == ref


def test_sparse_input_stays_sparse(monkeypatch, sandbag_api):
    """scipy-sparse input (any format) reaches the library as CSR and gives the dense answer."""
    import scipy.sparse as sp
    from pypairs_b200 import _ffi, pairs, utils
    _emulated_sandbag(monkeypatch)
    a = sandbag_api
    x = np.array(a["x"])
    sn, gn, lab = a["sample_names"], a["gene_names"], np.array(a["labels"])
    ann = {k: np.where(lab == i)[0].tolist() for i, k in enumerate(["G1", "S", "G2M"])}
    exp = pairs.sandbag(x, ann, gn, sn, fraction=0.65)
    for m in (sp.csr_matrix(x), sp.csc_matrix(x), sp.coo_matrix(x)):
        raw, _, _ = utils.parse_data(m, gn, sn)
        assert sp.issparse(raw) and raw.format == "csr"
        assert list(pairs.sandbag(m, ann, gn, sn, fraction=0.65).items()) == list(exp.items())
    keep, st, n_rows, n_cols, dev = _ffi.csr_arg(sp.csc_matrix(x.astype(np.int32)))
    assert dev is None
    assert (n_rows, n_cols) == x.shape and st.data_dtype == _ffi.PP_F64 and st.is_device == 0
    dup = sp.csr_matrix((np.array([1.0, 2.0, 5.0]), np.array([3, 3, 1]), np.array([0, 3] + [3] * (x.shape[0] - 1))),
                        shape=x.shape)
    keep, st, _, _, _ = _ffi.csr_arg(dup)  # duplicates are summed, as .toarray() would
    assert keep[0].tolist() == [5.0, 3.0] and keep[1].tolist() == [1, 3]
    # torch sparse-CSR: a repeated column inside a row is coalesced (summed) too, sorted or not
    import torch
    for cols in ([3, 3, 1], [3, 1, 3]):
        t = torch.sparse_csr_tensor(torch.tensor([0, 3] + [3] * (x.shape[0] - 1)), torch.tensor(cols),
                                    torch.tensor([1.0, 2.0, 5.0]), size=x.shape)
        keep, st, _, _, _ = _ffi.csr_arg(t)
        got = dict(zip(keep[2].tolist(), keep[0].tolist()))
        assert got == ({1: 5.0, 3: 3.0} if cols == [3, 3, 1] else {1: 2.0, 3: 6.0})
    t = torch.sparse_csr_tensor(torch.tensor([0, 2, 3] + [3] * (x.shape[0] - 2)), torch.tensor([4, 2, 2]),
                                torch.tensor([1.0, 2.0, 5.0]), size=x.shape)  # unsorted row, no duplicates: kept as is
    keep, st, _, _, _ = _ffi.csr_arg(t)
    assert keep[2].tolist() == [4, 2, 2] and keep[0].tolist() == [1.0, 2.0, 5.0]
    assert _ffi._pick_device(None, None) is None and _ffi._pick_device(1, None) == 1 and _ffi._pick_device(None, 1) == 1
    with pytest.raises(ValueError):
        _ffi._pick_device(0, 1)  # explicit device must agree with the tensor's device


def test_sandbag_input_errors(monkeypatch):
    from pypairs_b200 import pairs
    _emulated_sandbag(monkeypatch)
    x = np.arange(20.0).reshape(4, 5)
    with pytest.raises(ValueError):
        pairs.sandbag(x, {"a": [0, 1]})                                    # names required for ndarray
    with pytest.raises(ValueError):
        pairs.sandbag(x, None, list("abcde"), list("wxyz"))               # annotation required
    with pytest.raises(ValueError):
        pairs.sandbag(x, {"a": [0, 1], "b": [1, 2]}, list("abcde"), list("wxyz"))  # sample in two classes
    with pytest.raises(ValueError):
        pairs.sandbag(x, {"a": [0.5]}, list("abcde"), list("wxyz"))       # float annotation
    with pytest.raises(ValueError):
        pairs.sandbag("nope", {"a": [0]})


def test_to_boolean_mask_matches_reference_cases():
    """pypairs/tests/test_utils.py: int / str / None / [] / bool selections."""
    from pypairs_b200.utils import to_boolean_mask
    names = ["a", "b", "c", "d"]
    assert to_boolean_mask([0, 3], names).tolist() == [True, False, False, True]
    assert to_boolean_mask(["b", "d"], names).tolist() == [False, True, False, True]
    assert to_boolean_mask(None, names).tolist() == [True] * 4
    assert to_boolean_mask([], names).tolist() == [True] * 4
    assert to_boolean_mask([True, False, True, False], names).tolist() == [True, False, True, False]
    assert to_boolean_mask(2, names).tolist() == [False, False, True, False]
    with pytest.raises(ValueError):
        to_boolean_mask([0.5], names)


def test_filter_marker_pairs_vs_reference(cyclone_cases):
    from pypairs_b200.cyclone import filter_marker_pairs
    npz, meta = cyclone_cases
    for name, m in meta.items():
        markers = {k: [tuple(p) for p in v] for k, v in m["markers"].items()}
        pairs, used = filter_marker_pairs(markers, m["gene_names"])
        assert list(pairs.keys()) == m["classes"]
        for k in m["classes"]:
            assert np.array_equal(pairs[k], npz["{}_pairs_{}".format(name, k)])
            assert np.array_equal(used[k], npz["{}_used_{}".format(name, k)])


def test_cyclone_epilogue_rules():
    """max_class / 'N/A' and cc_prediction rules of cyclone.py:171-181 on hand-built score tables."""
    from pypairs_b200.cyclone import scores_to_frame
    sc = np.array([[0.1, 0.9, np.nan, 0.0, np.nan, 0.4, 0.5],
                   [0.2, 0.1, np.nan, 0.0, 0.3, 0.9, 0.5],
                   [0.7, 0.95, np.nan, 0.0, 0.2, 0.45, 0.5]])
    df = scores_to_frame(sc, ["G1", "S", "G2M"], list("abcdefg"))
    assert list(df["max_class"]) == ["G2M", "G2M", "N/A", "N/A", "S", "S", "G1"]
    assert [str(v) for v in df["cc_prediction"]] == ["G2M", "G2M", "nan", "S", "S", "S", "G1"]
    exp = orc.cyclone_epilogue({"G1": sc[0], "S": sc[1], "G2M": sc[2]}, list("abcdefg"))
    assert list(exp["max_class"]) == list(df["max_class"])
    assert [str(v) for v in exp["cc_prediction"]] == [str(v) for v in df["cc_prediction"]]
    df2 = scores_to_frame(sc[:2], ["alpha", "beta"], list("abcdefg"))
    assert list(df2.columns) == ["alpha", "beta", "max_class"]


@pytest.mark.parametrize("keys", [["G1", "S", "G2M"], ["G2M", "G1", "S"], ["a", "b"], ["G1", "S", "G2M", "M"]])
@pytest.mark.parametrize("arrow", [True, False])
def test_cyclone_epilogue_random_tables(monkeypatch, keys, arrow):
    """The vectorised epilogue (Arrow-decoded label columns, or the generic fallback) against the row-by-row
    restatement of cyclone.py:169-181, on tables with NaN columns / rows, zero rows and exact ties."""
    import pandas as pd
    from pypairs_b200 import cyclone as cy
    if not arrow:
        monkeypatch.setattr(pd, "StringDtype", type("NoStringDtype", (), {}))  # isinstance() fails -> fallback branch
    rng = np.random.default_rng(len(keys) * 2 + arrow)
    n = 400
    sc = rng.random((len(keys), n)).round(1)
    sc[:, :20] = np.nan
    sc[0, 20:40] = np.nan
    sc[:, 40:60] = 0.0
    sc[:, 60:90] = 0.4
    if len(keys) > 2:
        sc[[0, 2], 90:110] = np.nan
    sn = ["s{}".format(i) for i in range(n)]
    got = cy.scores_to_frame(sc, keys, sn)
    exp = orc.cyclone_epilogue({k: sc[i] for i, k in enumerate(keys)}, sn)
    assert list(got.columns) == list(exp.columns) and list(got.index) == sn
    for col in exp.columns:
        if col in keys:
            assert np.array_equal(got[col].to_numpy(), exp[col].to_numpy(), equal_nan=True)
        else:
            assert [str(v) for v in got[col]] == [str(v) for v in exp[col]], col


def test_quantile_transform_restrictions():
    """The sklearn pass-through is fitted on all cells: refused for rank-local shards and CUDA tensors."""
    from pypairs_b200 import cyclone as cyc
    x = np.random.default_rng(3).poisson(2.0, (30, 8)).astype(np.float64)
    with pytest.raises(ValueError):
        cyc._quantile_transform(x, True)
    from sklearn.preprocessing import QuantileTransformer
    import scipy.sparse as sp
    want = QuantileTransformer().fit_transform(x)
    assert np.array_equal(cyc._quantile_transform(x, False), want)
    assert np.array_equal(cyc._quantile_transform(sp.csr_matrix(x), False), want)


def test_default_marker_fixture(golden_dir):
    from pypairs_b200 import datasets
    d = datasets.default_cc_marker()
    assert {k: len(v) for k, v in d.items()} == {"G1": 2575, "S": 4097, "G2M": 1470}
    assert len({g for v in d.values() for p in v for g in p}) == 1479
    assert d["G1"][:3] == [["PSMC6", "PTEN"], ["PSMC6", "TOP2A"], ["PSMC6", "MDM2"]]
    if os.path.exists("/root/reference/pypairs/datasets/leng15_default_markers.json"):
        assert d == json.load(open("/root/reference/pypairs/datasets/leng15_default_markers.json"))


def test_first_positions_and_names_index():
    import pandas as pd
    from pypairs_b200 import utils
    from pypairs_b200.cyclone import _names_index
    rng = np.random.default_rng(3)
    for n_classes in (1, 3, 16, 17, 300):
        cls = rng.integers(0, n_classes, 5000)
        cls = cls[cls != n_classes // 2] if n_classes > 1 else cls      # one class never occurs
        exp = np.array([np.flatnonzero(cls == k)[0] if (cls == k).any() else len(cls) for k in range(n_classes)])
        assert np.array_equal(utils.first_positions(cls, n_classes), exp)
    assert np.array_equal(utils.first_positions(np.empty(0, dtype=np.int64), 3), [0, 0, 0])
    names = ["cell{}".format(i) for i in range(5000)]
    a, b = _names_index(names), pd.Index(names)
    assert a.equals(b) and a.dtype == b.dtype
    mixed = names[:2500] + [7] + names[2501:]                          # a non-str element: the generic conversion decides
    assert _names_index(mixed).equals(pd.Index(mixed))
    assert _names_index(b) is b


def test_cell_shard_partition():
    from pypairs_b200._dist import cell_shard
    for n in (0, 1, 7, 64, 1000003):
        for ws in (1, 2, 3, 8):
            spans = [cell_shard(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


def test_two_rank_gloo_merge():
    """world_size 2 on CPU (gloo): sharded sandbag merge and sharded cyclone gather give the 1-rank answer.
    The kernels are emulated by the oracle inside the worker (tests only)."""
    port = 29500 + os.getpid() % 2000
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE="2")
    procs = []
    for r in range(2):
        e = dict(env, RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_gloo_worker.py")], env=e,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, cwd=ROOT))
    outs = [p.communicate(timeout=300)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert all("GLOO_OK" in o for o in outs), "\n".join(outs)


@pytest.mark.parametrize("ws", [2, 3])
def test_nccl_worker_scenarios_emulated(ws):
    """The multi-rank GPU test script (tests/_nccl_worker.py) over gloo with oracle-backed kernels: every rank must
    issue the same sequence of collectives in every scenario (empty shards, gather_to, local names, NaN, probes)."""
    port = 31500 + os.getpid() % 2000 + ws
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(ws), PP_WORKER_EMULATE="1",
               PP_WORKER_WATCHDOG="280")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_nccl_worker.py")], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, cwd=ROOT) for r in range(ws)]
    outs = [p.communicate(timeout=400)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(o[-3000:] for o in outs)
    assert all("NCCL_OK" in o for o in outs)


def test_evaluate_prediction_vs_reference(golden_dir):
    """pypairs/tests/test_utils.py:7-30 known answers + tables produced by the reference (sklearn)."""
    from pypairs_b200 import utils
    pred = list("AABBCCDD")
    e = utils.evaluate_prediction(prediction=pred, reference=list("AABBCCDD"))
    assert np.array_equal(np.array(e.values, dtype=float), np.ones((5, 4)))
    e = utils.evaluate_prediction(prediction=pred, reference=list("DDBBCCAA"))
    assert np.array_equal(e.values, np.array([[0] * 4, [1] * 4, [1] * 4, [0] * 4, [0.5] * 4], dtype=float))
    for case in json.load(open(os.path.join(golden_dir, "evaluate_prediction.json"))):
        e = utils.evaluate_prediction(case["prediction"], case["reference"])
        assert [str(i) for i in e.index] == case["index"] and [str(c) for c in e.columns] == case["columns"]
        assert np.allclose(e.values.astype(float), np.array(case["values"]), rtol=0, atol=1e-15)


def test_marker_export_load_roundtrip(tmp_path, monkeypatch):
    """pypairs/tests/test_utils.py:33-58: default path under settings.writedir, failures are logged not raised."""
    from pypairs_b200 import datasets, settings, utils
    m = datasets.default_cc_marker()
    monkeypatch.setattr(settings, "writedir", str(tmp_path) + "/w/")
    monkeypatch.setattr(settings, "verbosity", 0)
    bad = "/not/existing/path/marker_exp.json"
    utils.export_marker(m, bad, defaultpath=False)
    assert not os.path.isfile(bad)
    utils.export_marker(m, "marker_exp.json")
    assert os.path.isfile(settings.writedir + "marker_exp.json")
    assert utils.load_marker(bad, defaultpath=False) is None
    assert utils.same_marker(m, utils.load_marker("marker_exp.json"))


def test_marker_table_and_cli(monkeypatch, tmp_path, sandbag_api):
    """Columnar MarkerTable == dict result; it feeds filter_marker_pairs like the dict; CLIs run end to end
    (kernels emulated by the oracle)."""
    from click.testing import CliRunner
    from pandas import DataFrame
    from pypairs_b200 import _ffi, pairs, utils
    from pypairs_b200.cli.cyclone.__main__ import main as cyclone_cli
    from pypairs_b200.cli.sandbag.__main__ import main as sandbag_cli
    from pypairs_b200.cyclone import filter_marker_pairs
    _emulated_sandbag(monkeypatch)
    a = sandbag_api
    x = np.array(a["x"])
    sn, gn, lab = a["sample_names"], a["gene_names"], np.array(a["labels"])
    ann = {k: np.where(lab == i)[0].tolist() for i, k in enumerate(["G1", "S", "G2M"])}
    d = pairs.sandbag(x, ann, gn, sn, fraction=0.65)
    t = pairs.sandbag(x, ann, gn, sn, fraction=0.65, as_table=True)
    assert isinstance(t, utils.MarkerTable) and list(t.to_dict().items()) == list(d.items())
    assert t.keys() == list(d.keys()) and t.counts() == {k: len(v) for k, v in d.items()} and len(t) == sum(t.counts().values())
    names = gn[5:] + ["other"]
    p_d, u_d = filter_marker_pairs(d, names)
    p_t, u_t = filter_marker_pairs(t, names)
    assert list(p_d.keys()) == list(p_t.keys())
    assert all(np.array_equal(p_d[k], p_t[k]) and np.array_equal(u_d[k], u_t[k]) for k in p_d)

    # CLI: sandbag
    data_csv, ann_csv, out_json = tmp_path / "d.csv", tmp_path / "a.csv", tmp_path / "m.json"
    DataFrame(x, index=sn, columns=gn).to_csv(data_csv)
    with open(ann_csv, "w") as f:
        f.write("name,class\n" + "".join("{},{}\n".format(sn[i], ["G1", "S", "G2M"][lab[i]]) for i in range(len(sn)) if lab[i] >= 0))
    r = CliRunner().invoke(sandbag_cli, ["-d", str(data_csv), "-a", str(ann_csv), "-f", "0.65", "-o", str(out_json), "-v", "0"])
    assert r.exit_code == 0, r.output
    got = json.load(open(out_json))
    assert utils.same_marker(got, d)

    # CLI: cyclone (kernel emulated)
    def fake_cyclone(xm, used_idx_list, pairs_list, iterations, min_iter, min_pairs, rng="mt19937", seed=2803,
                     row_begin=0, n_cells=None, **kw):
        xm = np.asarray(xm, dtype=np.float64)
        out = []
        for u, p in zip(used_idx_list, pairs_list):
            used = np.zeros(xm.shape[1], dtype=bool)
            used[u] = True
            out.append(orc.phase_scores(xm, iterations, min_iter, min_pairs, p, used, seed=seed))
        return np.stack(out), {}

    monkeypatch.setattr(_ffi, "cyclone", fake_cyclone)
    out_csv = tmp_path / "s.csv"
    r = CliRunner().invoke(cyclone_cli, ["-d", str(data_csv), "-m", str(out_json), "-i", "40", "-x", "10", "-y", "2",
                                         "-o", str(out_csv), "-v", "0"])
    assert r.exit_code == 0, r.output
    import pandas as pd
    sc = pd.read_csv(out_csv, index_col=0)
    exp = pairs.cyclone(DataFrame(x, index=sn, columns=gn), got, iterations=40, min_iter=10, min_pairs=2)
    assert list(sc.columns) == list(exp.columns) and np.allclose(sc[list(got.keys())].values, exp[list(got.keys())].values)
