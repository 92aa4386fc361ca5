"""Drop-in mirror of the reference seam ``myDetect.mfilter_coverage(moptions);
myDetect.mtest2(moptions)`` (bin/scripts/myDetect.py:639-641; also called from
mySimulate.py:243-245, mySimulat2.py:163-165, myDownSampling0.py:115-116).

Same function names, same ``moptions`` keys read and written, same table text -- but the
per-position tests and the neighbour combination run on the GPU.  Differences, all deliberate:
  * a position whose pooled values are all identical does not abort the run (scipy-1.2.1
    ``mannwhitneyu`` raises ValueError there, uncaught at :331): its U p-value is NaN;
  * the down-sampling branch (:345-361) draws from the library's own seeded counter-based
    stream (the reference uses numpy's unseeded global generator), ``moptions['seed']``;
  * ``moptions['_detector']`` may hold a ``Detector`` to reuse (else one is made on device 0).
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from .detect import DetectOptions, Detector, SignTestTable
from .pileup import Pileup


def mfilter_coverage(moptions: Dict) -> None:
    """myDetect.py:301-314.  Pure dictionary bookkeeping, kept on the host so that callers that
    inspect ``moptions[ds]`` afterwards see what the reference would leave there.  (The GPU
    applies the same filter again on the packed pileup; it is idempotent.)"""
    for dsn in moptions["ds2"]:
        curds = moptions[dsn]["norm_mean"]
        for sk in sorted(curds.keys()):
            for pk in sorted(curds[sk].keys()):
                if len(curds[sk][pk]) < moptions["MinCoverage"]:
                    del curds[sk][pk]
                    del moptions[dsn]["base"][sk][pk]
            if len(curds[sk]) == 0:
                del curds[sk]
                del moptions[dsn]["base"][sk]


def _detector(moptions: Dict) -> Detector:
    det = moptions.get("_detector")
    if det is None:
        det = Detector(int(moptions.get("device", 0)))
        moptions["_detector"] = det
    return det


OUTPUT_ERROR = 3  # myCom.py:6


def _run(moptions: Dict) -> SignTestTable:
    opt = DetectOptions.from_moptions(moptions)
    ds0, ds1 = moptions[moptions["ds2"][0]], moptions[moptions["ds2"][1]]
    pileup = Pileup.from_dicts(ds0, ds1)
    if moptions.get("outLevel", 2) <= OUTPUT_ERROR:  # the diagnostic of myDetect.py:432-434
        for sk, pk, b1, b0 in pileup.base_mismatch:
            extra = []
            for ds in (ds1, ds0):
                bd = ds.get("basedict")
                extra.append(list(bd[sk][pk].items()) if bd is not None else [])
            print("Error not equal", sk, pk, b1, b0, extra[0], extra[1])
    return _detector(moptions).detect(pileup, opt)


def save_test(moptions: Dict) -> None:
    """myDetect.py:522-538 (the table of the last mtest2 call)."""
    table: SignTestTable = moptions["_sign_test_table"]
    table.options.SaveTest = moptions["SaveTest"]
    table.save_test()


def mtest2(moptions: Dict) -> None:
    """myDetect.py:416-520: fills ``moptions['sign_test']`` and
    ``moptions['sorted_sign_test']`` and writes the table when SaveTest is set."""
    table = _run(moptions)
    moptions["_sign_test_table"] = table
    moptions["sign_test"] = table.to_sign_test()
    if moptions.get("mstd", 0):  # means / stds come from the device's moments (:437-438)
        moptions["sign_test_mstd"] = table.sign_test_mstd()
    if moptions.get("SaveTest", 0):
        save_test(moptions)
    st = moptions["sign_test"]
    if moptions.get("RegionRankbyST", 0) == 0:
        moptions["sorted_sign_test"] = [st[int(r)] for r in _detector(moptions).rank(table)]
    else:  # :463-515; the reference widens moptions['window'] by one as a side effect (:465)
        moptions["sorted_sign_test"] = [st[int(r)] for r in table.region_ranked()]
        moptions["window"] = moptions["window"] + 1


def getKStest(moptions: Dict, a, b, m_str: str) -> List:
    """myDetect.py:327-343 for one position: [(U,pU),(t,pt),(D,pks)] (clamped)."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    pl = Pileup.from_arrays(a, np.array([0, len(a)]), b, np.array([0, len(b)]), np.array([0]),
                            seg_names=[("x", m_str)])
    opt = DetectOptions(MinCoverage=3, testMethod="ks")
    t = _detector(moptions).detect(pl, opt)
    if len(t) == 0:
        raise ValueError("getKStest needs at least 3 values per group on the GPU path")
    return [(float(t.u_stat[0]), float(t.u_p[0])), (float(t.t_stat[0]), float(t.t_p[0])),
            (float(t.ks_d[0]), float(t.ks_p[0]))]


def called_sites(moptions: Dict):
    """The top-N site list mboxplot would plot (myDetect.py:279-297 with :153-164)."""
    table: SignTestTable = moptions["_sign_test_table"]
    table.options.topN = moptions["topN"]
    return table.called_sites()
