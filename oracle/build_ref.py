"""TEST INFRASTRUCTURE — recipe that makes the UNMODIFIED reference model files available where /root/reference is not
(the GPU box): the reference is pure Python, so "building" it is staging the files the hot path consists of, byte for
byte, from where they lie under /root/reference into oracle/_ref/ (git-ignored, so they never enter this repo's
history; NOT gpurun-ignored, so they travel to the GPU box like a built .so).  Nothing is edited; `sha256` of every
staged file is recorded in oracle/_ref/MANIFEST.txt and checked by tests/test_oracle_vs_reference.py when both trees
are present.

    python oracle/build_ref.py          # run by __graft_entry__.build() when /root/reference exists

Consumers (tests/, bench.py's reference arm, smoke) go through oracle/ref_loader.py, which prefers /root/reference and
falls back to oracle/_ref.  Nothing under aria_b200/ may touch either.
"""
import hashlib
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = os.environ.get("ARIA_REFERENCE_ROOT", "/root/reference")
DST_ROOT = os.path.join(HERE, "_ref")
FILES = [
    "aria/model/moe_lm.py", "aria/model/vision_encoder.py", "aria/model/projector.py",
    "aria/model/configuration_aria.py", "aria/model/modeling_aria.py", "aria/lora/layers.py",
]


def stage() -> bool:
    if not os.path.isfile(os.path.join(SRC_ROOT, FILES[0])):
        return False
    lines = []
    for rel in FILES:
        src, dst = os.path.join(SRC_ROOT, rel), os.path.join(DST_ROOT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        lines.append(f"{hashlib.sha256(open(dst, 'rb').read()).hexdigest()}  {rel}")
    with open(os.path.join(DST_ROOT, "MANIFEST.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    return True


if __name__ == "__main__":
    print("staged" if stage() else f"no reference tree under {SRC_ROOT}; nothing staged")
