#!/usr/bin/env python3
"""Measured FP32 CUDA-core peak of cuda:0 (scalar FFMA and packed FFMA2), the denominator of the FP32 roofline fractions (run under gpurun)."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from gato_b200 import native  # noqa: E402

print(json.dumps({"fp32_tflops_ffma": native.measure_fp32_peak(0, False), "fp32_tflops_ffma2": native.measure_fp32_peak(0, True)}))
