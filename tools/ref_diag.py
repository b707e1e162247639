"""Diagnostic (development aid): does the unmodified reference (oracle/_ref) finish on this box, and how fast?
   python tools/ref_diag.py [MB]"""
import os
import resource
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gen


def main():
    mb = float(sys.argv[1]) if len(sys.argv) > 1 else 48.0
    print("cpus", os.cpu_count(), "nofile", resource.getrlimit(resource.RLIMIT_NOFILE), "nproc",
          resource.getrlimit(resource.RLIMIT_NPROC), flush=True)
    tmp = tempfile.mkdtemp(prefix="refdiag_")
    path = os.path.join(tmp, "c.txt")
    n = gen.text_to_file(path, 1234, int(mb * 1e6 / 99.94), V=1000000)
    print("corpus", n, "bytes", flush=True)
    for procs, raise_nofile in ((0, False), (0, True), (32, True), (8, True)):
        env = dict(os.environ)
        env["DAMPR_REF_RAISE_NOFILE"] = "1" if raise_nofile else "0"
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_tfidf.py"), path, os.path.join(tmp, "idfs"),
                                str(procs)], cwd=tmp, capture_output=True, text=True, timeout=45, env=env)
            print("procs=%s raise_nofile=%s rc=%d %.1fs %.1f MB/s stderr=%r" % (
                procs or "all", raise_nofile, r.returncode, time.time() - t0, n / 1e6 / (time.time() - t0), r.stderr[-600:]), flush=True)
        except subprocess.TimeoutExpired as e:
            err = e.stderr.decode("utf-8", "replace")[-1500:] if e.stderr else ""
            print("procs=%s raise_nofile=%s TIMEOUT after 45s stderr=%r" % (procs or "all", raise_nofile, err), flush=True)
            pass


if __name__ == "__main__":
    main()
