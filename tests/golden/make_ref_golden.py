"""Generate tests/golden/ref_gpu_digests.json: SHA-256 digests of the outputs of the REFERENCE's own
kernels (oracle/_ref/libgiddy_ref.so, rebuilt for sm_100a by oracle/build_ref.sh) on every test case
for which the reference is well defined.  Must run on a GPU box:

    python tests/golden/make_ref_golden.py [output.json]

The inputs are regenerated from seeds by tests/cases.py, so only digests are committed.  Cases where
the reference kernels disagree with the oracle are reported (and left out), never silently accepted.
"""
import json
import os
import sys
import traceback

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import cases
    import libgiddy_b200 as g
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "ref_gpu_digests.json")
    reference = g.Library(os.path.join(ROOT, "oracle", "_ref", "libgiddy_ref.so"))
    digests, disagreements, failures = {}, {}, {}
    for case in cases.all_cases():
        if not case.ref_ok:
            continue
        try:
            got = cases.run_case_on_gpu(case, reference)
        except BaseException as error:  # a reference kernel may fault; a sticky CUDA error ends the run
            failures[case.name] = "%s: %s" % (type(error).__name__, str(error)[:200])
            traceback.print_exc()
            if "CUDA" in str(error) or "cuda" in str(error):
                break
            continue
        if (got == case.expected).all():
            digests[case.name] = cases.digest(got)
        else:
            bad = int((got != case.expected).sum())
            disagreements[case.name] = "%d of %d elements differ from the oracle" % (bad, len(got))
        print("%-70s %s" % (case.name, "ok" if case.name in digests else disagreements.get(case.name)), flush=True)
    result = {
        "what": "sha256 of the outputs of the reference's own kernels (eyalroz/libgiddy @ 74d3543, rebuilt for sm_100a) per tests/cases.py case",
        "device": torch.cuda.get_device_name(0), "digests": digests, "reference_disagrees_with_oracle": disagreements,
        "reference_failed": failures,
    }
    with open(out_path, "w") as f:
        json.dump(result, f, indent=1, sort_keys=True)
    print("%d digests, %d disagreements, %d failures -> %s" % (len(digests), len(disagreements), len(failures), out_path))


if __name__ == "__main__":
    main()
