"""SM clock / throttle-reason sampler for bench.py (NVML polled from a thread).

A config-2 step lasts tens of microseconds, so the whole timed loop is a few
milliseconds: `nvidia-smi -lms` cannot sample it.  NVML is polled in-process
every ~0.5 ms instead; the timed loop spends its time inside ctypes / CUDA calls
that release the GIL.
"""
from __future__ import annotations

import threading
import time

import numpy as np


class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap"}

    def __init__(self, torch_device, period_s=0.0005):
        self.period_s = period_s
        self.samples = []
        self.reason_bits = 0
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        self._h = None
        self._nv = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(torch_device).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            try:
                self._h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self._h = pynvml.nvmlDeviceGetHandleByIndex(torch_device.index or 0)
            self._nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._h = None

    def _poll(self):
        nv, h = self._nv, self._h
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                try:
                    self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            except Exception:
                pass
            time.sleep(self.period_s)

    def start(self):
        if self._h is None:
            return
        self._thread = threading.Thread(target=self._poll, daemon=True)
        self._thread.start()

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0,
               "how": "NVML polled every ~%g ms from a thread during the timed loops" % (self.period_s * 1e3)}
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
        if self.samples:
            out["sm_mhz"] = float(np.median(self.samples))
            out["samples"] = len(self.samples)
        out["reasons"] = [name for bit, name in self.REASONS.items() if self.reason_bits & bit]
        return out
