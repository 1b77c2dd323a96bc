"""Development aid (test infrastructure, uses the oracle): histogram of the EPA polytope sizes on a settled mixed pile.
Builds an instrumented copy of the oracle under /tmp (two counters added to epa()), steps a 16 384-body config-3 pile until it has
settled and prints the distribution of the final point / triangle counts per EPA run. The small shared-memory polytope of
npEpaSmallKernel (k_narrow_shapes.cuh: 10 points, 30 triangles, 32 edges) was sized from this output:

    points  cum: 4: .143  5: .363  6: .633  7: .821  8: .927  9: .971  10: .986  12: .996  24: 1
    tris    cum: 4: .143  8: .363  12: .628  16: .809  20: .917  24: .963  32: .990  48: .998  97 = max
"""
import ctypes as C
import os
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def build_instrumented(tmp="/tmp/b200_epa_hist"):
    os.makedirs(tmp, exist_ok=True)
    for f in os.listdir(os.path.join(ROOT, "oracle")):
        if f.endswith((".cpp", ".h")):
            shutil.copy(os.path.join(ROOT, "oracle", f), tmp)
    p = os.path.join(tmp, "om_narrow.cpp")
    s = open(p).read()
    head = "static void epa(const Simplex& g, const Support& A, const Support& B, EpaResult& out, NarrowStats& st)\n{"
    assert head in s
    s = s.replace(head, "uint32_t g_histP[64], g_histT[1100];\nextern \"C\" uint32_t* epa_hist_points() { return g_histP; }\n"
                        "extern \"C\" uint32_t* epa_hist_tris() { return g_histT; }\n" + head)
    s = s.replace("    if (status == 1) ++st.epaOutOfMemory;", "    g_histP[P.nP]++; g_histT[P.nT]++;\n    if (status == 1) ++st.epaOutOfMemory;")
    open(p, "w").write(s)
    for f in os.listdir(tmp):
        if f.endswith((".cpp", ".h")):
            q = os.path.join(tmp, f)
            text = open(q).read().replace('"../include/b200phys.h"', '"%s/include/b200phys.h"' % ROOT)
            open(q, "w").write(text)
    so = os.path.join(tmp, "liboracle.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", so] +
                          [os.path.join(tmp, f) for f in ("om_narrow.cpp", "om_step.cpp", "om_joints.cpp", "om_mass.cpp", "oracle_c.cpp")])
    return so


if __name__ == "__main__":
    import oracle_lib
    oracle_lib.LIB = build_instrumented()
    from d3d12renderer_b200 import host
    sc = host.build_mixed_slab(32, 16, 32, 0, 32)
    o = oracle_lib.Oracle(sc)
    lib = oracle_lib.load()
    lib.epa_hist_points.restype = C.POINTER(C.c_uint32); lib.epa_hist_tris.restype = C.POINTER(C.c_uint32)
    for _ in range(12):
        o.step(1 / 120., 10)
    hp, ht = lib.epa_hist_points(), lib.epa_hist_tris()
    for i in range(64): hp[i] = 0
    for i in range(1100): ht[i] = 0
    o.step(1 / 120., 10)
    a = np.array([hp[i] for i in range(64)]); t = np.array([ht[i] for i in range(1100)])
    print("EPA runs in one step:", int(a.sum()))
    print("points:", {i: int(a[i]) for i in range(64) if a[i]})
    cs = np.cumsum(a) / a.sum(); print("points cum:", {i: round(float(cs[i]), 3) for i in range(4, 25)})
    ct = np.cumsum(t) / t.sum(); print("tris cum:", {i: round(float(ct[i]), 3) for i in (4, 8, 12, 16, 20, 24, 28, 30, 32, 40, 48, 64, 96, 128, 160)})
    print(o.narrow_stats())
