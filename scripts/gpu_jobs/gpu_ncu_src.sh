# source-level ncu capture (stall samples per SASS line) of the particle kernels at cfg 1; prints the P(k) numbers first
set -x
python - <<'PY' > gpurun_out/pk_numbers.txt 2>&1
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import test_gpu_power_spectrum as T
from cafproject_b200.cube import CubeGPU
from cafproject_b200.power import cic_delta, cross_power
from cafproject_b200.run import cafcube
from cafproject_b200.synthetic_ic import make_ic
from cafproject_b200.timestep import Cosmology, TimeStepper
from oracle import cube_oracle as co
fk, ck = np.load("tests/golden/fk_table.npy"), np.load("tests/golden/ck_table.npy")
nc, nnt = 24, 2
states, sig, info = make_ic(nn=1, nc=nc, nnt=nnt, np_nc=2, seed=49, disp_rms=0.5)
O = co.Oracle(nn=1, nnt=nnt, nc=nc, np_nc=2, fk_table=fk, ck_table=ck)
O.load(states, sig); O.buffer_density(); O.buffer_x(); O.buffer_v()
tso = co.TimeStepper(co.Cosmology(), [0.0])
fo = T.oracle_run(O, tso, co)
G = CubeGPU(nc, nnt, fk, ck, np_nc=2, tanf_lut=co.tanf_lut())
G.particle_initialization(states[0], sig); G.buffer_density(); G.buffer_x(); G.buffer_v()
got = {}
ts = TimeStepper(Cosmology(), [0.0])
n = cafcube(G, ts, on_checkpoint=lambda z, st, s: got.update(state=st))
xi = cross_power(cic_delta([got["state"]], 1, nc, nnt), cic_delta(fo, 1, nc, nnt), 200.0)
nyq = 4 * nc // 2
k = np.arange(1, xi.shape[1] + 1)
print("steps gpu/oracle", n, tso.istep)
print("k  P_gpu/P_oracle-1  r")
for i in range(len(k)):
    if k[i] < nyq / 2: print(k[i], "%.2e" % (xi[2][i] / xi[3][i] - 1), "%.6f" % xi[7][i])
PY
cat gpurun_out/pk_numbers.txt | tail -30
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:k_fine_deposit|k_drift_place_p|k_fine_kick_p|k_fft_z_green|k_fft_x_inv|k_fft_y" -c 7 -f -o gpurun_out/prof_src \
    python bench.py --nc 128 --nnt 2 --steps 1 --warmup 3 --no-cpu --no-e2e --profile > gpurun_out/ncu_src.log 2>&1
echo "ncu rc=$?"
