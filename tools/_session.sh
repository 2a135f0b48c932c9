mkdir -p gpurun_out
nvidia-smi -L
( time timeout 600 python -m pytest tests/test_gpu_ensemble.py -m gpu -x -q -k "nccl" ) > gpurun_out/r2h_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2h_pytest.log
tail -30 gpurun_out/r2h_pytest.log
# single-process two-GPU ensemble through qp_ens_create (ncclCommInitAll)
timeout 300 python - <<'PY' 2>&1 | tail -5
import numpy as np, sys
sys.path.insert(0, '.')
import qprop_b200 as qp
from qprop_b200.ensemble import LibraryEnsemble, EnsembleChebyPropagator
comm = LibraryEnsemble.local([0, 1])
print('transport', comm.transport, comm.world, comm.n_local)
B = 13
w = qp.workloads.config3_transmon(n_sites=4, levels=4, B=B, nt=6, dt=0.5)
H0, H1, H2 = w["ops"]
bound = float((abs(H0) + 0.1 * abs(H1) + 0.1 * abs(H2)).sum(axis=1).max())
members = [EnsembleChebyPropagator(w["ops"], w["controls"], w["scales"], w["psi0"], w["tlist"], -bound, bound, comm.contexts[i], rank=i, world=2) for i in range(2)]
for m in members: m.propagate()
states = comm.gather_states([m.state for m in members], B)
ref = EnsembleChebyPropagator(w["ops"], w["controls"], w["scales"], w["psi0"], w["tlist"], -bound, bound, comm.contexts[0])
ref.propagate()
full = ref.state.to_host().reshape(256, B)
print('single-process 2-GPU gather rel err', np.linalg.norm(states - full) / np.linalg.norm(full))
vals = comm.gather_expvals([np.asarray(m.state.norm()).reshape(1, -1) for m in members], B)
print('norms', np.max(np.abs(vals.real - 1)))
print('SINGLE_PROCESS_OK' if np.linalg.norm(states - full) < 1e-13 else 'SINGLE_PROCESS_FAILED')
PY
