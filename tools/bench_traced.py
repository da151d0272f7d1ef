"""bench.py with a watchdog: prints the Python stack of every thread if the run is still going after
BENCH_WATCHDOG seconds (default 150) -- for runs under ncu that do not come back."""
import faulthandler
import os
import runpy
import sys

faulthandler.dump_traceback_later(int(os.environ.get('BENCH_WATCHDOG', '150')), exit=True)
sys.argv = [os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'bench.py')] + sys.argv[1:]
runpy.run_path(sys.argv[0], run_name='__main__')
