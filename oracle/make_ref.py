"""Vendoring recipe for the reference arm — TEST / BENCH INFRASTRUCTURE, not product code.

The reference (Roestlab/diffusion-deconvolution-dia-msms-data) is pure Python: there is nothing to compile.  This
script copies the UNMODIFIED modules of the hot path from /root/reference into oracle/_ref/ (git-ignored, NOT
gpurun-ignored: like a built .so it travels to the GPU box, where /root/reference does not exist) together with the
two import shims the reference needs in this image (rotary_embedding_torch: restated third-party arithmetic,
SURVEY.md §8c; duckdb: import-only, the .npy path never calls it).  `bench.py --impl reference` then times the
reference's own `DDIMDiffusionModel._train_one_batch` on the host cores (`cpu_baseline.kind = "reference"`), and
`bench.py` reports the same code on the B200 in eager fp32 as `gpu_eager_baseline`.

    python oracle/make_ref.py     # (re)creates oracle/_ref/; a no-op with exit 0 when /root/reference is absent
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/dquartic"
DST = os.path.join(HERE, "_ref")
FILES = ["model/model.py", "model/model_interface.py", "model/unet1d.py", "model/building_blocks.py",
         "utils/data_loader.py"]


def main():
    if not os.path.isdir(SRC):
        print(f"make_ref: {SRC} not present (GPU box) - keeping whatever oracle/_ref holds")
        return 0
    pkg = os.path.join(DST, "dquartic")
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    for rel in FILES:
        dst = os.path.join(pkg, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
    # package markers written here (the reference's own __init__ imports cli / data_generation, which need polars etc.)
    for d in ("", "model", "utils"):
        open(os.path.join(pkg, d, "__init__.py"), "w").close()
    for shim in ("rotary_embedding_torch.py", "duckdb.py"):
        shutil.copyfile(os.path.join(HERE, "_shims", shim), os.path.join(DST, shim))
    print(f"make_ref: vendored {len(FILES)} reference modules into {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
