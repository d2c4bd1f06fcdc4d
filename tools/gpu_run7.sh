cd $GRAFT_REPO_ROOT
for pv in 0 2; do
DV_GEMM_PAIR=$pv timeout 600 python - <<'PY'
import numpy as np, os, sys
sys.path.insert(0, os.getcwd())
from tests.test_parity_exact_gpu import _synthetic_pair
import bench
from d_vins_b200 import capi
from oracle import quant, weights
W = weights.synth_all()
e = capi.Engine(height=480, width=752, weights_path=bench.make_weights())
for (M,N) in [(300,400),(150,662),(512,512)]:
    k0,k1,d0,d1=_synthetic_pair(M,N,M+N)
    mq,sq=quant.lightglue_q(weights.sub(W,"lg."),k0,k1,d0,d1,480,752,480,752)
    for rep in range(3):
        mg,sg=e.lg_match(k0,k1,d0,d1,480,752,480,752)
        print(os.environ.get("DV_GEMM_PAIR"), M,N,rep,len(mg),len(mq),np.array_equal(mg,mq), np.abs(np.log(sg)-np.log(sq)).max() if len(mg)==len(mq) else None)
PY
done
