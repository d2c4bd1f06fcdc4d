cd $GRAFT_REPO_ROOT
for cfg in "0 0" "2 0" "0 1" "2 1"; do
set -- $cfg
DV_GEMM_PAIR=$1 DV_ATTN_PERSIST=$2 timeout 600 python - <<'PY'
import numpy as np, os, sys
sys.path.insert(0, os.getcwd())
from tests.test_parity_exact_gpu import _synthetic_pair
import bench
from d_vins_b200 import capi
from oracle import quant, weights
W = weights.synth_all()
e = capi.Engine(height=480, width=752, weights_path=bench.make_weights())
for (M,N) in [(1024,1024),(37,1000)]:
    k0,k1,d0,d1=_synthetic_pair(M,N,M+N)
    mq,sq=quant.lightglue_q(weights.sub(W,"lg."),k0,k1,d0,d1,480,752,480,752)
    mg,sg=e.lg_match(k0,k1,d0,d1,480,752,480,752)
    sq_=set(map(tuple,mq)); sg_=set(map(tuple,mg))
    print(os.environ.get("DV_GEMM_PAIR"), os.environ.get("DV_ATTN_PERSIST"), M,N,len(mg),len(mq),np.array_equal(mg,mq), sorted(sq_-sg_), sorted(sg_-sq_))
    if sq_-sg_:
        i=[k for k,m in enumerate(mq) if tuple(m) in (sq_-sg_)]
        print("  oracle scores of missing:", sq[i])
PY
done
