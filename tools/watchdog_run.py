#!/usr/bin/env python
"""Run a script under a host-side watchdog: after --after seconds every thread's Python stack goes to stderr and the
process exits (so a hung kernel costs seconds of GPU time, and the stack says which call never returned).
    python tools/watchdog_run.py --after 120 bench.py --steps 20"""
import faulthandler
import runpy
import sys

after = 120
argv = sys.argv[1:]
if argv and argv[0] == "--after":
    after = int(argv[1])
    argv = argv[2:]
faulthandler.dump_traceback_later(after, exit=True)
sys.argv = argv
runpy.run_path(argv[0], run_name="__main__")
