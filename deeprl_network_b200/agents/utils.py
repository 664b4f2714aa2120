"""Host-side helper of the agent API: the learning-rate schedule (``lr_decay`` in MODEL_CONFIG; the reference keeps
it next to its buffers, agents/utils.py:917-930).  Every ``get(n_step)`` first advances the consumed-step count and
then evaluates the schedule, so the first update already runs slightly below ``lr_init``."""


class Scheduler:
    LINEAR, CONSTANT = 'linear', 'constant'

    def __init__(self, val_init, val_min=0, total_step=0, decay='linear'):
        self.val, self.val_min, self.decay = val_init, val_min, decay
        self.N = float(total_step)        # horizon of the linear ramp, in environment steps
        self.n = 0                        # environment steps consumed so far

    def _at(self, n):
        if self.decay != self.LINEAR:
            return self.val
        ramp = self.val * (1 - n / self.N)
        return ramp if ramp > self.val_min else self.val_min

    def get(self, n_step):
        self.n += n_step
        return self._at(self.n)
