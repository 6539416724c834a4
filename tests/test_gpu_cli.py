"""The two console scripts (pypairs/cli/sandbag/__main__.py, pypairs/cli/cyclone/__main__.py) end to end on the GPU:
CSV in, marker JSON / score CSV out, as real subprocesses (`python -m pypairs_b200.cli.*`), checked against the oracle."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle as orc
from test_gpu_sandbag import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(module, *args):
    r = subprocess.run([sys.executable, "-m", module] + [str(a) for a in args], cwd=ROOT, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r


def test_sandbag_and_cyclone_cli(tmp_path):
    import pandas as pd
    rng = np.random.default_rng(77)
    x, lab = synth(rng, 90, 120, 3, frac_de=0.2)
    x[:, 5] = 0  # an unexpressed gene: dropped like in the reference (utils.py:368)
    sn = ["s{}".format(i) for i in range(90)]
    gn = ["g{}".format(i) for i in range(120)]
    cls = ["G1", "S", "G2M"]
    data_csv, ann_csv, marker_json, score_csv = (tmp_path / n for n in ("d.csv", "a.csv", "m.json", "s.csv"))
    pd.DataFrame(x, index=sn, columns=gn).to_csv(data_csv)
    with open(ann_csv, "w") as f:
        f.write("name,class\n" + "".join("{},{}\n".format(sn[i], cls[lab[i]]) for i in range(90)))
    _run("pypairs_b200.cli.sandbag", "-d", data_csv, "-a", ann_csv, "-f", 0.6, "-o", marker_json, "-v", 0)
    got = json.load(open(marker_json))
    # oracle: classes in annotation-file order of first appearance, expressed genes only
    order = []
    for i in range(90):
        if cls[lab[i]] not in order:
            order.append(cls[lab[i]])
    cats = np.array([order.index(cls[k]) for k in lab])
    expressed = (x != 0).any(axis=0)
    o = np.argsort(cats, kind="stable")
    rr, cc, kk = orc.sandbag_arrays(x[o][:, expressed], cats[o], 3, 0.6)
    gk = np.array(gn)[expressed]
    exp = {}
    for r, c, k in zip(rr, cc, kk):
        exp.setdefault(order[k], []).append([gk[r], gk[c]])
    assert list(got.keys()) == list(exp.keys()) and got == exp and sum(map(len, got.values())) > 20

    _run("pypairs_b200.cli.cyclone", "-d", data_csv, "-m", marker_json, "-i", 60, "-x", 10, "-y", 2, "-o", score_csv, "-v", 0)
    sc = pd.read_csv(score_csv, index_col=0)
    assert list(sc.index) == sn and "max_class" in sc.columns
    pr, us = orc.filter_marker_pairs(got, gn)
    for k in got:
        ref = orc.phase_scores(x, 60, 10, 2, pr[k], us[k])
        assert np.allclose(sc[k].values, ref, rtol=0, atol=1e-12, equal_nan=True)  # CSV round trip of the doubles
    # no -o: the result goes to stdout
    r = _run("pypairs_b200.cli.cyclone", "-d", data_csv, "-m", marker_json, "-i", 20, "-x", 5, "-y", 2, "-v", 0)
    assert "max_class" in r.stdout
