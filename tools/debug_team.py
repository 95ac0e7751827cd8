"""A/B of the two K2 mappings on a contact scene, one model step at a time from the SAME state (the thread kernel's):
    python tools/debug_team.py c3|c4|c5 [K]
prints the worst rollout per step and its state rows under both kernels."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import torch

from scenes import boxer_setup, push_setup
from mppi_isaac_b200.backend import CudaBackend

DEV = "cuda:0"
which = sys.argv[1] if len(sys.argv) > 1 else "c3"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
if which == "c3":
    T = 20
    sc, p, s0 = boxer_setup(K=K, T=T)
    rng = np.random.default_rng(0)
    actions = np.stack([rng.uniform(0.3, 1.2, (T, K)), rng.uniform(-1.0, 1.0, (T, K))], axis=1).astype(np.float32)
    free_rows = [1]
elif which == "c4":
    T = 25
    sc, p, s0 = push_setup(K=K, T=T, noise=True, block_pos=(0.62, 1.5, 0.1))
    actions = np.random.default_rng(3).uniform(-0.6, 0.6, (T, 3, K)).astype(np.float32)
    actions[:, 0] = 0.5 + 0.1 * actions[:, 0]
    free_rows = [1]
else:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
    from test_gpu_sizes import _pick_scene
    T = 30
    sc, p, s0 = _pick_scene(K, T)
    rng = np.random.default_rng(5)
    actions = rng.uniform(-0.2, 0.2, (T, sc.nu, K)).astype(np.float32)
    actions[:, 7:9] = -0.15 + 0.05 * actions[:, 7:9]
    free_rows = [sc.model.free_actor[f] for f in range(sc.model.nfree)]


def backend(team):
    os.environ["MPPIB_K2_TEAM"] = "1" if team else "0"
    be = CudaBackend(DEV)
    be.create(sc.model, p)
    return be


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).to(DEV)


ref, team = backend(False), backend(True)
a_d, root_d = dev(actions), dev(sc.root_state0)
NS, nd2 = ref.state_size(), 2 * sc.ndof
state = np.zeros((NS, K), np.float32)
state[:nd2] = s0[:, None]
for f, actor in enumerate(free_rows):
    state[nd2 + 13 * f: nd2 + 13 * (f + 1)] = sc.root_state0[actor][:, None]
obs_r = torch.zeros((ref.obs_size(), T, K), device=DEV)
obs_t = torch.zeros((ref.obs_size(), T, K), device=DEV)
np.set_printoptions(precision=5, suppress=True, linewidth=200)
for t in range(T):
    sr, st = dev(state), dev(state)
    ref.rollout(None, sr, a_d, t, 1, obs_r, root0=root_d)
    team.rollout(None, st, a_d, t, 1, obs_t, root0=root_d)
    torch.cuda.synchronize()
    r, g = sr.cpu().numpy(), st.cpu().numpy()
    err = np.abs(r - g).max(axis=0)
    k = int(err.argmax())
    oerr = (obs_r[:, t] - obs_t[:, t]).abs().max().item() if t > 0 else 0.0
    print(f"step {t}: worst |state| diff {err.max():.3e} at rollout {k}; {(err > 1e-3).sum()} rollouts above 1e-3; obs diff {oerr:.3e}")
    if err.max() > 1e-3:
        print("  before", state[:, k])
        print("  thread", r[:, k])
        print("  team  ", g[:, k])
        print("  action", actions[t, :, k])
    state = r
