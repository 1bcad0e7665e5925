#!/bin/bash
mkdir -p gpurun_out
HPS_EARLY_ZERO=0 HPS_TIMELINE_REPS=2 timeout 200 python scripts/step_timeline.py > gpurun_out/e_timeline_noearly.txt 2> gpurun_out/e_err1.txt
HPS_EARLY_ZERO=0 HPS_TAIL_PRIORITY=-1 HPS_TIMELINE_REPS=2 timeout 200 python scripts/step_timeline.py > gpurun_out/e_timeline_noearly_tailprio.txt 2> gpurun_out/e_err2.txt
HPS_EARLY_ZERO=1 HPS_TAIL_PRIORITY=-1 HPS_TIMELINE_REPS=2 timeout 200 python scripts/step_timeline.py > gpurun_out/e_timeline_early_tailprio.txt 2> gpurun_out/e_err3.txt
tail -2 gpurun_out/e_err1.txt
grep span gpurun_out/e_timeline_*.txt
