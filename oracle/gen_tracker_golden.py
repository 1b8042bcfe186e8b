"""Golden vectors for MetricTracker (TEST INFRASTRUCTURE ONLY): runs the reference's own
src/metrics/tracker.py:4-72 (importable here: it only needs pandas) on a fixed update sequence and stores
the sequence with the results in tests/golden/metric_tracker.json.

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_tracker_golden.py
"""
import importlib.util
import json
import os
import random

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_tracker", "/root/reference/src/metrics/tracker.py")
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)


class RefTracker(mod.MetricTracker):
    """The reference class with `reset` restated: its `self._data[col].values[:] = 0` needs a writable view, which
    the pandas in this image (>= 3, copy-on-write) no longer hands out.  update / avg / result / keys are the
    reference's own code."""

    def reset(self):
        for col in self._data.columns:
            self._data[col] = 0.0


rng = random.Random(7)
keys = ["SISNRMetric", "SISNRiMetric", "loss"]
cases = []
for case in range(3):
    tr = RefTracker(*keys, writer=None)
    seq = []
    for _ in range(40):
        k = rng.choice(keys)
        v = rng.uniform(-30.0, 30.0)
        n = rng.choice([1, 1, 1, 2, 5])
        tr.update(k, v, n=n)
        seq.append([k, v, n])
    mid = {k: float(tr.avg(k)) for k in keys}
    res = {k: float(v) for k, v in tr.result().items()}
    tr.reset()
    after_reset = {k: float(v) for k, v in tr.result().items()}
    tr.update("loss", 1.5)
    after_one = {k: float(v) for k, v in tr.result().items()}
    cases.append({"keys": keys, "updates": seq, "avg": mid, "result": res, "after_reset": after_reset,
                  "after_reset_then_loss_1.5": after_one, "keys_listed": list(tr.keys())})
out = os.path.join(HERE, "..", "tests", "golden", "metric_tracker.json")
with open(out, "w") as f:
    json.dump({"source": "src/metrics/tracker.py (reference), pandas " + __import__("pandas").__version__, "cases": cases}, f)
print("wrote", out)
