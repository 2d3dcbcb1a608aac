"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):

    compute-sanitizer --tool memcheck python tools/sanitize_cases.py

Runs the first 3000 reads of the bundled Ec10k input, a 2000-read synthetic both-strand input
(k=31, 100 bp) and every micro case of tests/cases.py through the C ABI on cuda:0, including the
exports and a second cb_string_graph (cut-log restore).  No oracle: the sanitizer is the checker."""
import gzip
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from cases import all_cases
    from cloudbrush_b200 import BrushConfig, BrushGpu, synth
    work = []
    sfa = gzip.open(os.path.join(ROOT, "tests", "golden", "ec10k_reads.sfa.gz")).read()
    work.append(("ec10k_3000", b"\n".join(sfa.split(b"\n")[:3000]) + b"\n", 21, 36))
    buf, m = synth.make_config("cfg3_50k", scale_genome=4000)
    work.append(("cfg3_4k", bytes(buf), m["k"], m["readlen"]))
    quick = os.environ.get("CB_SANITIZE_QUICK")
    for name, (b, k, L) in sorted(all_cases().items()):
        if quick and name not in ("branch", "long_group", "many_branches", "line_formats"):
            continue
        work.append((name, b, k, L))
    for name, b, k, L in work:
        g = BrushGpu(BrushConfig(k=k, readlen=L))
        g.load_sfa_buffer(b)
        g.run()
        n = [len(g.edges(ck)) for ck in (1, 2, 3)]
        g.string_graph()
        g.nodes()
        print(name, "nodes", g.counter("nodes"), "edges", n, flush=True)
        g.close()


if __name__ == "__main__":
    main()
