"""tools/pump_cost.py -- what ONE map input costs in Python on its way through the pump, measured without a GPU: the
real InputPreprocessor / InputPumper and bench.py's null control plane around a hash context that returns at once.
Two figures: wall time per input (best / median of R passes; pin the process to one core for stable numbers) and,
deterministic, the bytecode instructions and Python frames per input counted with sys.monitoring, broken down by
function with --by-function.

    taskset -c 5 python tools/pump_cost.py [--by-function] [R]"""
import collections
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench
from modal_client_b200 import _backend, blob_utils


class InstantContext:
    device = -1

    def hash_buffers(self, bufs, flags=3):
        n = len(bufs)
        return np.zeros((n, 32), np.uint8), np.zeros((n, 16), np.uint8), np.zeros(n, np.uint64)


_backend.set_context(InstantContext())
_backend.context_pool = lambda k: [InstantContext() for _ in range(k)]
blob_utils._upload_to_s3_url = bench._null_put
args = [a for a in sys.argv[1:] if not a.startswith("--")]
R = int(args[0]) if args else 15
N = 100_000
payload = bytes(262144)


def one_pass(n):
    stub = bench.NullStub()
    t0 = time.perf_counter()
    bench.run_map_pump([payload] * n, stub)  # includes filling the raw-input queue, as the bench's timed region does
    return time.perf_counter() - t0


one_pass(N)
walls = sorted(one_pass(N) for _ in range(R))
print(f"wall per input over {N} inputs: best {walls[0] / N * 1e6:.2f} us, median {walls[R // 2] / N * 1e6:.2f} us ({R} passes)")

mon = sys.monitoring
TOOL = 2
mon.use_tool_id(TOOL, "pump_cost")
instr, frames = collections.Counter(), collections.Counter()
mon.register_callback(TOOL, mon.events.INSTRUCTION, lambda code, off: instr.update((code,)))
mon.register_callback(TOOL, mon.events.PY_START, lambda code, off: frames.update((code,)))
M = 16384
mon.set_events(TOOL, mon.events.INSTRUCTION | mon.events.PY_START)
one_pass(M)
mon.set_events(TOOL, 0)
print(f"bytecode instructions per input: {sum(instr.values()) / M:.1f}; Python frames entered per input: {sum(frames.values()) / M:.2f}")
if "--by-function" in sys.argv:
    for code, c in instr.most_common(25):
        print(f"{c / M:8.1f} instr  {frames[code] / M:5.2f} frames  {os.path.basename(code.co_filename)}:{code.co_firstlineno} {code.co_name}")
