for v in "" ${VARIANTS:-nostore}; do
  if [ -n "$v" ]; then export POINTDSC_B200_LIB=$PWD/tools/bin/lib_$v.so; else unset POINTDSC_B200_LIB; fi
  python - <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from pointdsc_b200 import PointDSC
from pointdsc_b200.synth import make_batch
z = np.load("tests/golden/snapshot_3dmatch.npz"); sd = {k: torch.from_numpy(z[k]) for k in z.files}
m = PointDSC(num_layers=12); m.load_state_dict(sd, strict=False); m = m.cuda().eval()
base = make_batch(range(16), 1000, "3dmatch", 0.3)
cp, s, t = (base[k].repeat(16, 1, 1).cuda() for k in ("corr_pos", "src_keypts", "tgt_keypts"))
for _ in range(3): m.run(cp, s, t)
m.profile(True)
for _ in range(5): m.run(cp, s, t)
p = m.profile_read()
print(os.environ.get("POINTDSC_B200_LIB", "product"), {k: round(v[0] / 5, 3) for k, v in p.items()})
PY
done
