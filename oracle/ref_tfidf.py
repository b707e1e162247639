"""Runs the UNMODIFIED reference on its own TF-IDF benchmark — TEST / BASELINE INFRASTRUCTURE ONLY.

The reference package is the pip install of /root/reference under oracle/_ref/ (oracle/make_ref.sh).
The statements below are the body of the reference's benchmarks/tf-idf-dampr.py:9-21 (the script itself
is not part of the installed package), with the input path, the sink directory and the run name taken
from argv instead of being hard-coded to /tmp/idfs; everything else — chunking by cpu count, the
tokenising lambda, count / cross_right / sink_tsv, the stock MTRunner with the reference's default
settings — is the reference's public API and stock code path.

    python oracle/ref_tfidf.py <corpus> <out_dir> [max_processes]

Run as a separate process (bench.py --impl reference, the bench's parity check): this repository ships its
own `dampr` package for drop-in use, so the reference must be imported from oracle/_ref in a process whose
sys.path does not contain the repository root.
"""
import math
import multiprocessing
import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def main(argv):
    path, out = argv[1], argv[2]
    try:
        # environment, not reference code: the reference's reduce stage opens one run file per map task and
        # partition (dataset.py:571-579); with ~100 host cores that exceeds the usual soft limit of 1024 open
        # files, a worker dies with EMFILE and the reference's parent then waits forever
        import resource
        soft, hard = resource.getrlimit(resource.RLIMIT_NOFILE)
        if os.environ.get("DAMPR_REF_RAISE_NOFILE", "1") == "1" and (hard == resource.RLIM_INFINITY or soft < hard):
            resource.setrlimit(resource.RLIMIT_NOFILE, (hard if hard != resource.RLIM_INFINITY else 1 << 20, hard))
    except Exception:
        pass
    sys.path[:] = [REF] + [p for p in sys.path if os.path.abspath(p or ".") not in (os.path.dirname(HERE), HERE)]
    import dampr
    assert os.path.abspath(dampr.__file__).startswith(REF), "not the reference: %s" % dampr.__file__
    from dampr import Dampr, settings
    if len(argv) > 3 and int(argv[3]) > 0:
        settings.max_processes = int(argv[3])
    name = "dampr_ref_%d" % os.getpid()

    chunk_size = os.stat(path).st_size / multiprocessing.cpu_count()
    docs = Dampr.text(path, chunk_size + 1)

    RX = re.compile(r'[^\w]+')
    doc_freq = docs \
        .flat_map(lambda x: set(RX.split(x.lower()))) \
        .count(reduce_buffer=float('inf'))

    idf = doc_freq.cross_right(docs.len(),
                               lambda df, total: (df[0], df[1], math.log(1 + (float(total) / df[1]))),
                               memory=True)

    idf.sink_tsv(out).run(name=name)
    # the reference never removes its stage directories under /tmp
    shutil.rmtree(os.path.join("/tmp", name), ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
