"""CPU oracle: n-step returns / advantages, on-policy buffer protocol, LR schedule.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  PINNED against the unmodified
reference buffers through ``tests/golden/buffer_*.npz`` and ``scheduler.npz``.

Follows ``/root/reference/agents/utils.py``:
  * ``_add_R_Adv``   (global reward)        -- :763-775 (IA2C) / :837-855 (MA2C)
  * ``_add_s_R_Adv`` (spatial discount)     -- :800-816 (IA2C) / :888-912 (MA2C)
  * pre-step vs post-step dones, fp32 casts -- :731-761, :823-835
  * ``Scheduler``                           -- :917-930
"""
import numpy as np


def nstep_returns(r, v, done_post, R_end, gamma, alpha=-1.0, dist=None):
    """Vector restatement of the reverse scans.

    r         [T, N] float64  (normalised rewards; for alpha<0 every column equal)
    v         [T, N] float64  rollout values
    done_post [T]    bool     done AFTER step t
    R_end     [N]    float64  bootstrap values
    returns Rs, Advs [N, T] float32 (the reference casts at the end, :830-831)

    alpha < 0 :  R <- r_t + gamma * R * (1 - done)
    alpha > 0 :  R <- gamma * R * (1 - done);  R += sum_{d=0..maxdist_i} alpha**d * sum_{j: dist(i,j)=d} r_tj
                 (accumulated in ascending d like the reference inner loop)
    """
    r = np.asarray(r, dtype=np.float64)
    v = np.asarray(v, dtype=np.float64)
    T, N = v.shape
    Rs = np.zeros((N, T))
    Advs = np.zeros((N, T))
    for i in range(N):
        R = float(R_end[i])
        if alpha > 0:
            dm = np.asarray(dist[i])
            maxd = int(dm.max())
        for t in range(T - 1, -1, -1):
            nd = 1.0 - float(done_post[t])
            if alpha < 0:
                R = r[t, i] + gamma * R * nd
            else:
                R = gamma * R * nd
                for d in range(maxd + 1):
                    R += (alpha ** d) * np.sum(r[t][dm == d])
            Rs[i, t] = R
            Advs[i, t] = R - v[t, i]
    return Rs.astype(np.float32), Advs.astype(np.float32)


class RolloutBuffer:
    """List-append buffer with the reference's done bookkeeping
    (``dones = [prev_done] + post-step dones``; sample returns the PRE-step dones).
    Multi-agent layout like ``MultiAgentOnPolicyBuffer`` (agents/utils.py:819-835)."""

    def __init__(self, gamma, alpha, dist):
        self.gamma, self.alpha, self.dist = gamma, alpha, dist
        self.reset(False)

    def reset(self, done=False):
        self.obs, self.adds, self.acts, self.rs, self.vs = [], [], [], [], []
        self.dones = [done]

    def add_transition(self, ob, p, a, r, v, done):
        self.obs.append(ob); self.adds.append(p); self.acts.append(a)
        self.rs.append(r); self.vs.append(v); self.dones.append(done)

    def finish(self, R_end):
        """Returns/advantages + pre-step dones; then reset(last done) like :833-834."""
        N = len(self.vs[0])
        r = np.array([np.broadcast_to(np.asarray(x, dtype=np.float64), (N,)) for x in self.rs])
        v = np.array(self.vs, dtype=np.float64)
        Rs, Advs = nstep_returns(r, v, self.dones[1:], R_end, self.gamma, self.alpha, self.dist)
        dones = np.array(self.dones[:-1], dtype=bool)
        self.reset(self.dones[-1])
        return dones, Rs, Advs

    def sample_transition(self, R_end):
        obs = np.transpose(np.array(self.obs, dtype=np.float32), (1, 0, 2))
        ps = np.transpose(np.array(self.adds, dtype=np.float32), (1, 0, 2))
        acts = np.transpose(np.array(self.acts, dtype=np.int32))
        dones, Rs, Advs = self.finish(R_end)
        return obs, ps, acts, dones, Rs, Advs


class Scheduler:
    """agents/utils.py:917-930"""

    def __init__(self, val_init, val_min=0, total_step=0, decay='linear'):
        self.val, self.N, self.val_min, self.decay, self.n = val_init, float(total_step), val_min, decay, 0

    def get(self, n_step):
        self.n += n_step
        if self.decay == 'linear':
            return max(self.val_min, self.val * (1 - self.n / self.N))
        return self.val
