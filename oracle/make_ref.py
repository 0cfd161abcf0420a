"""Recipe for oracle/_ref/: the UNMODIFIED reference sources of the hot path, copied from
/root/reference when it is mounted (the build container) into the git-ignored oracle/_ref/ so
that they travel to the GPU box with the repo snapshot (oracle/_ref/ is NOT in .gpurunignore).

TEST / BASELINE INFRASTRUCTURE ONLY.  Nothing under oracle/_ref/ is committed, shipped or
imported by eav_b200/; it is executed only by `bench.py --impl reference` (the CPU arm: stock
EEGNet_tor + Trainer_uni objects through oracle/ref_shim.py) and by the oracle pin tests.
The reference is pure Python, so "building" it is a file copy; nothing is compiled.

    python oracle/make_ref.py            # no-op (keeps an existing copy) when /root/reference is absent
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("EAV_REFERENCE_SRC", "/root/reference")
# the files SURVEY.md section 8(a) names for the path (+ the package marker the imports need)
FILES = ("Dataload_eeg.py", "EAV_datasplit.py", os.path.join("CNN_torch", "EEGNet_tor.py"),
         os.path.join("CNN_torch", "CNN_EEG.py"))


def make_ref(verbose=False) -> bool:
    """Returns True when oracle/_ref holds the reference files (fresh copy or an earlier one)."""
    have_src = all(os.path.isfile(os.path.join(SRC, f)) for f in FILES)
    if have_src:
        for f in FILES:
            dst = os.path.join(DEST, f)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(os.path.join(SRC, f), dst)
        with open(os.path.join(DEST, "PROVENANCE.txt"), "w") as fh:
            fh.write(f"verbatim copies of {', '.join(FILES)} from {SRC} (nubcico/EAV), made by oracle/make_ref.py\n")
        if verbose:
            print(f"oracle/_ref: copied {len(FILES)} files from {SRC}")
    return all(os.path.isfile(os.path.join(DEST, f)) for f in FILES)


if __name__ == "__main__":
    ok = make_ref(verbose=True)
    print("oracle/_ref ready" if ok else "oracle/_ref NOT available (no reference mounted, no earlier copy)")
    sys.exit(0)
