"""Environment layer.  `PBO_Env` is the reference's two-line delegation
(`src/environment/basic_environment.py:6-22`); `BatchedEnv` is its B-lane counterpart whose `step` takes
a device tensor of actions and never leaves the GPU."""
from typing import Any


class PBO_Env:
    """An environment with a problem and an optimizer (same surface as the reference)."""

    def __init__(self, problem, optimizer):
        self.problem = problem
        self.optimizer = optimizer

    def reset(self):
        self.problem.reset()
        return self.optimizer.init_population(self.problem)

    def step(self, action: Any):
        return self.optimizer.update(action, self.problem)


class BatchedEnv:
    """B lanes = (instance, run) pairs advanced in lockstep by one engine (`*_Engine` in optimizer.py).
    reset() -> state tensor ; step(action tensor) -> (state, reward, done) tensors, all on the device."""

    def __init__(self, engine):
        self.optimizer = engine
        self.problem = engine.pack

    def reset(self):
        return self.optimizer.init_population()

    def step(self, action):
        return self.optimizer.update(action)
