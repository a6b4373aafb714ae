"""Debug aid: where does the reference CUDA run differ from the oracle, and is it repeatable?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import Oracle, RefEngine, configs
from tests.common import variant
cfg = sys.argv[1] if len(sys.argv) > 1 else "small435"
prm = configs.params(cfg)
left, right = configs.pair(prm, seed=3)
orc = Oracle()
ref = orc.pipeline(prm, left, right)
prm_nolr = variant(prm, lr_max_diff=255, mf_size=1)
runs = []
for it in range(4):
    r = RefEngine(prm_nolr)
    r.compute_host(left, right)
    runs.append((r.stage("leftDisp").copy(), r.stage("rightDisp").copy(), r.stage("LAll").copy()))
    r.close()
for it, (dl, dr, la) in enumerate(runs):
    ne = np.argwhere(dl != ref["disp_wta"])
    print(f"run {it}: leftDisp(no LR) vs oracle disp_wta: {len(ne)} differ; rightDisp differ {(dr != ref['disp_right']).sum()}; LAll differ {(la != ref['LAll']).sum()}")
    for y, x in ne[:8]:
        v = la[y, x].astype(int)
        m = v.min(); am = int(v.argmin())
        srt = np.sort(v)
        print(f"   ({y},{x}) ref={dl[y,x]!r} oracle={ref['disp_wta'][y,x]!r} argmin={am} min={m} vals around={v[max(am-2,0):am+3].tolist()} second={srt[1]} count_min={(v==m).sum()}")
print("runs identical:", all(np.array_equal(runs[0][0], r[0]) for r in runs))
