"""Poor man's sampling profiler for multi-threaded task processes (no py-spy in the image):

    python tools/stack_sampler.py OUT.txt script.py [args...]

runs ``script.py`` and samples every thread's Python stack every 5 ms (``sys._current_frames``); on exit / SIGTERM writes, per
thread name, the most frequent innermost frames and the most frequent "our code" frames (first frame inside this repository).
Used to find where a ps task of the control-plane tier spends its time (handler threads block in C calls that cProfile's
per-thread hooks never see)."""
import collections
import os
import runpy
import signal
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    out, script = sys.argv[1], sys.argv[2]
    sys.argv = sys.argv[2:]
    inner = collections.Counter()
    ours = collections.Counter()
    total = collections.Counter()
    stop = threading.Event()

    def sample():
        me = threading.get_ident()
        while not stop.is_set():
            names = {t.ident: t.name for t in threading.enumerate()}
            for tid, fr in sys._current_frames().items():
                if tid == me:
                    continue
                name = names.get(tid, "?").split("-")[0:3]
                name = "-".join(name)
                total[name] += 1
                c = fr.f_code
                inner[(name, "%s:%d %s" % (os.path.relpath(c.co_filename, ROOT) if c.co_filename.startswith(ROOT) else os.path.basename(c.co_filename), fr.f_lineno, c.co_name))] += 1
                f = fr
                while f is not None and not f.f_code.co_filename.startswith(os.path.join(ROOT, "distributed_tensorflow_b200")):
                    f = f.f_back
                if f is not None:
                    ours[(name, "%s:%d %s" % (os.path.relpath(f.f_code.co_filename, ROOT), f.f_lineno, f.f_code.co_name))] += 1
            time.sleep(0.005)

    def dump(*_):
        stop.set()
        with open(out, "w") as f:
            for title, ctr in (("innermost frame", inner), ("innermost frame inside the package", ours)):
                f.write("== %s (samples, thread, frame)\n" % title)
                for (name, where), n in ctr.most_common(60):
                    f.write("%6d  %5.1f%%  %-28s %s\n" % (n, 100.0 * n / max(1, total[name]), name, where))
        os._exit(0)
    signal.signal(signal.SIGTERM, dump)
    threading.Thread(target=sample, daemon=True, name="sampler").start()
    sys.path.insert(0, os.path.dirname(os.path.abspath(script)))
    try:
        runpy.run_path(script, run_name="__main__")
    finally:
        dump()


if __name__ == "__main__":
    main()
