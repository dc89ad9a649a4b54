"""Recipe for oracle/_ref/: makes the UNMODIFIED hot-path modules of the reference available next to the oracle, so that
bench.py's CPU arm (`--impl reference`, `cpu_baseline`) can time the reference's own code on the GPU box's host cores
(/root/reference does not exist there).

TEST / MEASUREMENT INFRASTRUCTURE ONLY: nothing under openobj_b200/ imports oracle/_ref.  The directory is git-ignored (it
never enters the history) but NOT gpurun-ignored, so it travels with the snapshot like the built .so files.  Only the nine
files of objnerf/ that SURVEY.md section 8(a) cites, plus the room_0 JSON, are taken -- byte for byte, from where they lie
under /root/reference.  Run by __graft_entry__.build() whenever /root/reference is present.

    python oracle/build_ref.py [--check]
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("OPENOBJ_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["objnerf/model.py", "objnerf/embedding.py", "objnerf/render_rays.py", "objnerf/loss.py", "objnerf/utils.py",
         "objnerf/cfg.py", "objnerf/vis.py", "objnerf/trainer.py", "objnerf/vmap.py", "objnerf/configs/Replica/room_0.json"]


def build(check=False):
    if not os.path.isfile(os.path.join(SRC, "objnerf", "vmap.py")):
        print("oracle/build_ref: %s not present; leaving %s as it is (%s)" %
              (SRC, DST, "present" if os.path.isdir(DST) else "absent"))
        return os.path.isdir(DST)
    ok = True
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        if check:
            same = os.path.isfile(dst) and filecmp.cmp(src, dst, shallow=False)
            ok &= same
            print("%-44s %s" % (rel, "identical" if same else "MISSING / DIFFERENT"))
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    if not check:
        with open(os.path.join(DST, "README"), "w") as f:
            f.write("Unmodified files of BIT-DYN/OpenObj objnerf/, placed here by oracle/build_ref.py for the CPU reference arm of\n"
                    "bench.py.  Not part of the repository (git-ignored); never imported by openobj_b200/.\n")
    return ok


if __name__ == "__main__":
    sys.exit(0 if build(check="--check" in sys.argv) else 1)
