cd $GRAFT_REPO_ROOT
cat > /tmp/exp.py <<'PY'
import sys, os
sys.path.insert(0, os.environ["GRAFT_REPO_ROOT"])
import numpy as np, torch
from pydem_b200 import synth, tile as T
n = 4096
E = synth.conditioned_fractal_dem(n, 0)
dt = T.DeviceTile(n, n, stream=torch.cuda.current_stream().cuda_stream)
dt.set_spacing(30.0, 30.0); dt.upload(T.F_ELEV, E)
for rep in range(3):
    dt.slopes_directions()
    st = dt.uca(drain_pits=1)
    print(st["ms_sweep"], st["ms_sweep_kernel"], st["n_drained"], st["n_undone"])
PY

for lib in libpydem_b200.so libpydem_b200_nofence.so; do echo "lib=$lib"; PYDEM_B200_LIB=$GRAFT_REPO_ROOT/pydem_b200/$lib PYDEM_B200_SWEEP=tile PYDEM_B200_TS_TILE=1 PYDEM_B200_TS_DEBUG=45 timeout 60 python /tmp/exp.py 2>&1 | grep -v "timeline\|CTA-time\|passes" | tail -4 | cut -c1-300; done
