"""Measure the FP32-FFMA and SFU peaks of cuda:0 (csrc/microbench.cu) and print them as JSON."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mae_b200 import peaks  # noqa: E402

out = peaks.measure(torch.device("cuda:0"), reps=10)
out["nominal"] = {"ffma_tflops": out["sm_count"] * 128 * 2 * 1.965e-3, "mufu_gops": out["sm_count"] * 16 * 1.965}
print(json.dumps(out, indent=1))
