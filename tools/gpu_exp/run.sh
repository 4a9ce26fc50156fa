(timeout 900 python -m pytest tests -m gpu -x -q -k "edt" 2>&1 | tail -5)
python - <<PY
import time, numpy as np, torch, sys
sys.path.insert(0,'.')
from tests import helpers as H
from f110_gymnasium_ros2_jazzy_b200 import BatchSim
from scipy.ndimage import distance_transform_edt
m=H.load('maps'); shape=tuple(int(v) for v in m['Shanghai_map__shape'])
free=np.unpackbits(m['Shanghai_map__bits'])[:shape[0]*shape[1]].reshape(shape)
sim=BatchSim(1,1)
sim.set_map_image(free,0.06505,[0,0,0])
t=time.perf_counter(); sim.set_map_image(free,0.06505,[0,0,0]); print('device EDT 2000x2000 incl upload: %.1f ms'%((time.perf_counter()-t)*1e3))
t=time.perf_counter(); d=0.06505*distance_transform_edt(np.where(free,255.,0.)); print('scipy EDT: %.1f ms'%((time.perf_counter()-t)*1e3))
PY
