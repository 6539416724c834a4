"""Generate tests/golden/* by running the UNMODIFIED reference (Numba) in this container.

    python oracle/make_golden.py

Needs /root/reference (not present on the GPU box -- which is why the outputs are
committed).  The reference imports `anndata` and `colorama`, neither installed here;
oracle/shims/ provides inert stand-ins.  `pairs.cyclone` itself is broken under
pandas >= 3 (pypairs/tools/cyclone.py:173 positional Series lookup), so cyclone is
driven at kernel level: filter_marker_pairs + get_phase_scores + get_proportion
(cyclone.py:201-257, 273-311).

Outputs
    sandbag_minref.json       reference output on its own KAT (tests/test_sandbag.py:26-59), fractions .65/.5
    sandbag_cases.npz         synthetic matrices + reference result lists (C in 1,2,3,10; several fractions)
    sandbag_api.json          full-API runs (DataFrame / ndarray, annotation int/str/bool, filters)
    cyclone_cases.npz/.json   synthetic matrices + marker dicts + reference scores / observed counters
    mt19937.json              generator known answers + permutation-table digests
    default_cc_marker.json    compact columnar form of the bundled leng15 marker set (data fixture)
    evaluate_prediction.json  utils.evaluate_prediction tables (sklearn) on four label vectors
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(0, "/root/reference")

import numba  # noqa: E402
import pypairs  # noqa: E402
from pandas import DataFrame  # noqa: E402
from pypairs import pairs, settings, utils  # noqa: E402
from pypairs.tools import cyclone as ref_cyclone  # noqa: E402
from pypairs.tools import sandbag as ref_sandbag  # noqa: E402

settings.verbosity = 0


def synth_counts(rng, n_cells, n_genes, n_classes, frac_de=0.3, skew=False):
    """Poisson counts with class-specific DE genes (SURVEY.md section 8d generator)."""
    if skew:
        p = 1.0 / (np.arange(n_classes) + 1.0)
        labels = rng.choice(n_classes, size=n_cells, p=p / p.sum())
    else:
        labels = rng.integers(0, n_classes, n_cells)
    for k in range(n_classes):  # no empty class
        labels[k % n_cells] = k
    base = rng.normal(0.0, 1.5, n_genes)
    de = rng.random(n_genes) < frac_de
    de_class = rng.integers(0, n_classes, n_genes)
    sign = rng.choice([-1.0, 1.0], n_genes)
    lam = np.tile(3.0 * np.exp(base), (n_cells, 1))
    for g in np.where(de)[0]:
        lam[labels == de_class[g], g] *= 2.0 ** (2.0 * sign[g])
    x = rng.poisson(lam).astype(np.float64)
    return x, labels.astype(np.int64)


def ref_check_pairs(x, cats, thr):
    settings.enable_jit = True
    settings.n_jobs = 4
    fn = utils.parallel_njit(ref_sandbag.check_pairs)
    return fn(np.ascontiguousarray(x, dtype=np.float64), cats.astype(np.int64), thr.astype(np.int64))


def main():
    os.makedirs(OUT, exist_ok=True)

    # ---------------------------------------------------------------- min_ref KAT through the public API
    mat = np.array([
        [48, 10, 63, 30, 1, 8, 4, 15, 20, 50, 20, 18], [0] * 12,
        [12, 42, 26, 8, 12, 1, 8, 38, 20, 15, 61, 52], [22, 8, 8, 4, 43, 33, 21, 3, 1, 6, 4, 6],
        [1, 41, 6, 5, 15, 12, 16, 8, 9, 3, 33, 29], [30, 4, 41, 38, 3, 3, 8, 1, 0, 51, 6, 2], [3] * 12,
        [2, 31, 6, 5, 31, 26, 29, 19, 28, 13, 29, 31], [38, 12, 12, 43, 0, 18, 1, 4, 6, 51, 8, 9],
        [10, 18, 18, 11, 6, 1, 5, 21, 5, 14, 18, 16], [99, 1, 87, 66, 18, 19, 11, 2, 19, 75, 2, 4]])
    cats = {"G1": [0, 2, 3, 8], "S": [4, 5, 6], "G2M": [1, 7, 9, 10]}
    snames = ["A_G1", "B_G2M", "C_G1", "D_G1", "E_S", "F_S", "G_S", "H_G2M", "I_G1", "J_G2M", "K_G2M", "L_X"]
    gnames = ["X{}".format(i) for i in range(11)]
    minref = {"matrix_genes_by_samples": mat.tolist(), "annotation": cats, "sample_names": snames,
              "gene_names": gnames, "results": {}}
    settings.enable_jit = True
    settings.n_jobs = 4
    for frac in (0.65, 0.5, 0.4, 1.0):
        res = pairs.sandbag(mat.T, cats, gnames, snames, fraction=frac)
        minref["results"][repr(frac)] = [[str(k), [[str(a), str(b)] for a, b in v]] for k, v in res.items()]
    json.dump(minref, open(os.path.join(OUT, "sandbag_minref.json"), "w"), indent=0)

    # ---------------------------------------------------------------- synthetic kernel-level sandbag cases
    rng = np.random.default_rng(20260)
    cases = {}
    spec = [  # name, cells, genes, classes, fraction, skew
        ("c3_f65", 96, 70, 3, 0.65, False), ("c3_f50", 61, 50, 3, 0.5, False), ("c3_f40", 45, 40, 3, 0.4, False),
        ("c2_f65", 77, 60, 2, 0.65, False), ("c2_f50", 40, 40, 2, 0.5, False), ("c1_f65", 33, 30, 1, 0.65, False),
        ("c1_f40", 33, 30, 1, 0.4, False), ("c10_f65", 300, 64, 10, 0.65, True), ("c10_f70", 130, 48, 10, 0.7, False),
        ("c3_f100", 50, 40, 3, 1.0, False), ("c3_tiny", 50, 40, 3, 1e-9, False), ("c4_skew", 129, 65, 4, 0.6, True),
    ]
    for name, n, g, c, frac, skew in spec:
        x, lab = synth_counts(rng, n, g, c, skew=skew)
        if name == "c3_f65":  # sign of zero and float32-colliding doubles must not merge/split ties
            x[:, 3] = np.where(x[:, 3] == 0, -0.0, x[:, 3])
            x[:, 5] = x[:, 4] + 1e-12
            x[:, 7] = 5.0  # constant gene
        sizes = np.bincount(lab, minlength=c)
        thr = ref_sandbag.calc_thresholds(np.array([lab == k for k in range(c)]), frac).astype(np.int64)
        res = ref_check_pairs(x, lab, thr)
        rows, cols = np.where(res != -1)
        cases[name + "_x"] = x
        cases[name + "_cats"] = lab
        cases[name + "_thr"] = thr
        cases[name + "_frac"] = np.float64(frac)
        cases[name + "_rows"] = rows.astype(np.int32)
        cases[name + "_cols"] = cols.astype(np.int32)
        cases[name + "_cls"] = res[rows, cols].astype(np.int32)
        print("sandbag", name, "sizes", sizes.tolist(), "thr", thr.tolist(), "markers", len(rows))
    np.savez_compressed(os.path.join(OUT, "sandbag_cases.npz"), **cases)

    # ---------------------------------------------------------------- full-API sandbag runs
    x, lab = synth_counts(rng, 60, 45, 3, frac_de=0.4)
    x[:, 9] = 0.0  # unexpressed gene -> filtered (utils.py:368)
    names3 = ["G1", "S", "G2M"]
    sn = ["cell{}".format(i) for i in range(60)]
    gn = ["gene{}".format(i) for i in range(45)]
    lab[57:] = -1  # uncategorised samples
    ann_int = {k: np.where(lab == i)[0].tolist() for i, k in enumerate(names3)}
    ann_str = {k: [sn[j] for j in np.where(lab == i)[0]] for i, k in enumerate(names3)}
    ann_bool = {k: (lab == i).tolist() for i, k in enumerate(names3)}
    api = {"x": x.tolist(), "sample_names": sn, "gene_names": gn, "labels": lab.tolist(), "runs": []}

    def rec(tag, res, **kw):
        api["runs"].append({"tag": tag, "kwargs": kw,
                            "result": [[str(k), [[str(a), str(b)] for a, b in v]] for k, v in res.items()]})

    rec("ndarray_int", pairs.sandbag(x, ann_int, gn, sn, fraction=0.65), annotation="int", fraction=0.65)
    rec("ndarray_str", pairs.sandbag(x, ann_str, gn, sn, fraction=0.65), annotation="str", fraction=0.65)
    rec("ndarray_bool", pairs.sandbag(x, ann_bool, gn, sn, fraction=0.65), annotation="bool", fraction=0.65)
    df = DataFrame(x, index=sn, columns=gn)
    rec("dataframe_int", pairs.sandbag(df, ann_int, fraction=0.6), annotation="int", fraction=0.6, dataframe=True)
    fg = list(range(0, 45, 2))
    fs = [sn[i] for i in range(5, 55)]
    rec("filters", pairs.sandbag(x, ann_int, gn, sn, fraction=0.6, filter_genes=fg, filter_samples=fs),
        annotation="int", fraction=0.6, filter_genes=fg, filter_samples=fs)
    ann_rev = {k: ann_int[k] for k in ["G2M", "G1", "S"]}  # dict order = class order (utils.py:249)
    rec("class_order", pairs.sandbag(x, ann_rev, gn, sn, fraction=0.65), annotation="int_rev", fraction=0.65)
    ann_empty = dict(ann_int)
    ann_empty["EMPTY"] = [58]  # class that loses all samples to filter_samples -> dropped (sandbag.py:202-209)
    rec("empty_class", pairs.sandbag(x, ann_empty, gn, sn, fraction=0.65, filter_samples=list(range(0, 57))),
        annotation="int_empty", fraction=0.65, filter_samples=list(range(0, 57)))
    json.dump(api, open(os.path.join(OUT, "sandbag_api.json"), "w"))

    # ---------------------------------------------------------------- cyclone kernel-level cases
    markers = json.load(open("/root/reference/pypairs/datasets/leng15_default_markers.json"))
    allgenes = sorted({g for v in markers.values() for p in v for g in p})
    compact = {"source": "pypairs/datasets/leng15_default_markers.json (BSD-3, Fechtner & Scialdone)",
               "genes": allgenes, "classes": {}}
    gi = {g: i for i, g in enumerate(allgenes)}
    for k, v in markers.items():
        compact["classes"][k] = [[gi[a] for a, _ in v], [gi[b] for _, b in v]]
    json.dump(compact, open(os.path.join(OUT, "default_cc_marker.json"), "w"), separators=(",", ":"))

    cyc_npz, cyc_meta = {}, {}
    rng = np.random.default_rng(20261)

    def run_cyclone(tag, x, gene_names, mp, iterations, min_iter, min_pairs):
        mpi, used = ref_cyclone.filter_marker_pairs(mp, gene_names)
        meta = {"gene_names": list(gene_names), "markers": {k: [list(p) for p in v] for k, v in mp.items()},
                "iterations": iterations, "min_iter": min_iter, "min_pairs": min_pairs, "classes": list(mp.keys())}
        cyc_npz[tag + "_x"] = x
        for k in mp:
            if len(mpi[k]) == 0:
                continue
            sc = ref_cyclone.get_phase_scores(x.astype(np.float64), iterations, min_iter, min_pairs, mpi[k], used[k])
            cyc_npz["{}_score_{}".format(tag, k)] = sc
            cyc_npz["{}_pairs_{}".format(tag, k)] = np.asarray(mpi[k], dtype=np.int32)
            cyc_npz["{}_used_{}".format(tag, k)] = used[k]
            print("cyclone", tag, k, "pairs", len(mpi[k]), "used", int(used[k].sum()), "score[:4]", sc[:4])
        cyc_meta[tag] = meta

    # (a) subset of default markers, tie-heavy Poisson, names shuffled, some marker genes absent
    sub = {k: [tuple(p) for p in v[:160]] for k, v in markers.items()}
    genes_a = sorted({g for v in sub.values() for p in v for g in p})
    drop = set(rng.choice(genes_a, 12, replace=False).tolist())
    gene_names = [g for g in genes_a if g not in drop] + ["FILL{}".format(i) for i in range(40)]
    rng.shuffle(gene_names)
    xa = rng.poisson(2.0, (40, len(gene_names))).astype(np.float64)
    xa[3] = xa[2]  # identical cells -> identical scores
    run_cyclone("a_poisson", xa, gene_names, sub, 1000, 100, 50)
    # (b) tie-free lognormal, reference-test settings min_iter=10, min_pairs=1 (tests/test_cyclone.py:24)
    xb = rng.lognormal(0.0, 1.0, (24, len(gene_names)))
    run_cyclone("b_lognormal", xb, gene_names, sub, 200, 10, 1)
    # (c) NaN / gate paths: min_pairs above any hit count -> 0.0 ; iterations < min_iter -> NaN
    run_cyclone("c_minpairs", xa[:8], gene_names, sub, 120, 100, 10000)
    run_cyclone("d_miniter", xa[:8], gene_names, sub, 50, 100, 1)
    # (e) duplicate pairs, (g,g) pairs, duplicate gene names (last wins), custom class names / order
    gn_e = ["g{}".format(i) for i in range(30)]
    gn_e[7] = "g3"  # duplicate name: index 7 wins for "g3"
    mp_e = {"zeta": [("g0", "g1"), ("g0", "g1"), ("g2", "g2"), ("g3", "g4"), ("g5", "missing"), ("g29", "g10")] +
            [("g{}".format(i), "g{}".format((i * 7 + 3) % 30)) for i in range(30)],
            "alpha": [("g{}".format((i * 5) % 30), "g{}".format((i * 11 + 1) % 30)) for i in range(45)]}
    xe = rng.poisson(4.0, (16, 30)).astype(np.float64)
    run_cyclone("e_quirks", xe, gn_e, mp_e, 300, 100, 5)
    np.savez_compressed(os.path.join(OUT, "cyclone_cases.npz"), **cyc_npz)
    json.dump(cyc_meta, open(os.path.join(OUT, "cyclone_cases.json"), "w"))

    # ---------------------------------------------------------------- MT19937 stream known answers from Numba
    @numba.njit
    def nb_tables(seed, n, iters):
        np.random.seed(seed)
        a = np.arange(n)
        out = np.empty((iters, n), dtype=np.int64)
        for k in range(iters):
            np.random.shuffle(a)
            out[k] = a
        return out

    @numba.njit
    def nb_raw(seed, n):
        np.random.seed(seed)
        out = np.empty(n, dtype=np.int64)
        for i in range(n):
            out[i] = np.random.randint(0, 4294967296)
        return out

    mt = {"seed": 2803, "numba": numba.__version__, "tables": {}}
    for n in (1195, 1199, 917, 5, 17, 64, 1479):
        t = nb_tables(2803, n, 1000)
        mt["tables"][str(n)] = {"sha256_int64le_1000xn": hashlib.sha256(t.astype("<i8").tobytes()).hexdigest(),
                                "first_row_head": t[0][:8].tolist(), "last_row_head": t[-1][:8].tolist()}
    mt["tables"]["7_seed99"] = {"rows": nb_tables(99, 7, 6).tolist()}
    json.dump(mt, open(os.path.join(OUT, "mt19937.json"), "w"), indent=0)
    # ---------------------------------------------------------------- utils.evaluate_prediction (pypairs/utils.py:23-94)
    ev_cases = [(list("AABBCCDD"), list("AABBCCDD")), (list("AABBCCDD"), list("DDBBCCAA")),
                (["G1", "S", "S", "G2M", "N/A", "G1", "G1"], ["G1", "S", "G2M", "G2M", "S", "G1", "S"]),
                (list("aabbccab"), list("abcabcab"))]
    ev = []
    for pred, ref in ev_cases:
        r = utils.evaluate_prediction(prediction=pred, reference=ref)
        ev.append({"prediction": pred, "reference": ref, "index": [str(i) for i in r.index],
                   "columns": [str(c) for c in r.columns], "values": r.values.astype(float).tolist()})
    json.dump(ev, open(os.path.join(OUT, "evaluate_prediction.json"), "w"))
    print("reference version", pypairs.__version__, "numba", numba.__version__)


if __name__ == "__main__":
    main()
