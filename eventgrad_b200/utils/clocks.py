"""Sample GPU clocks / throttle reasons with nvidia-smi during a timed region
(B200_PROFILING.md "clocks DURING the timed region")."""
from __future__ import annotations

import statistics
import subprocess
import threading
from typing import Dict, List

_Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
      "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
      "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")


class ClockSampler:
    def __init__(self, gpu_index: int = 0, period_ms: int = 100):
        self.gpu, self.period = gpu_index, period_ms
        self.proc = None
        self.lines: List[str] = []
        self._t = None

    def start(self) -> "ClockSampler":
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={_Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.gpu), "-lms", str(self.period)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except (FileNotFoundError, OSError):
            self.proc = None
            return self

        def pump():
            for ln in self.proc.stdout:
                self.lines.append(ln.strip())
        self._t = threading.Thread(target=pump, daemon=True)
        self._t.start()
        return self

    def stop(self) -> Dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw = [], [], []
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}
