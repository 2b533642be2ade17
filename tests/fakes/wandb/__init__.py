"""Stand-in for the `wandb` package (absent from the image): rsl_rl/runners/on_policy_runner.py imports it at module level and calls
wandb.log / wandb.Histogram from OnPolicyRunner.log.  Test scaffolding only."""
logged = []


class Histogram:
    def __init__(self, data=None, **kw):
        self.data = data


def log(d, step=None, **kw):
    logged.append((step, dict(d)))


def init(*a, **k):
    return None
