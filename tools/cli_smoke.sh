#!/bin/bash
# End-to-end smoke of the drop-in CLI on a GPU box: generate-config -> synthetic .npy pool -> `dquartic train` (2 epochs,
# checkpoints in the reference format) -> resume from the latest checkpoint.  Shrunken m/z axis (640) so checkpoints stay small.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
export PYTHONPATH=$ROOT/diffusion-deconvolution-dia-msms-data_b200:$PYTHONPATH
W=${1:-/tmp/dq_cli_smoke}
rm -rf $W && mkdir -p $W && cd $W
python -m dquartic.cli generate-config cfg.json
python - <<'PY'
import json, numpy as np, sys
sys.path.insert(0, ".")
from dquartic.utils.synthetic import synth_pool
ms2, ms1 = synth_pool(24, 34, 640, seed=7, density=0.2)
np.save("ms2.npy", ms2); np.save("ms1.npy", ms1)
c = json.load(open("cfg.json"))
c["model"]["UNet1d"]["downsample_dim"] = 640
c["model"].update(batch_size=8, num_epochs=2, warmup_epochs=1, checkpoint_path="ckpt/best_model.ckpt")
c["wandb"]["use_wandb"] = False
c["data"]["parquet_directory"] = None   # the generated default points at a parquet directory; NPY paths are exclusive with it
json.dump(c, open("cfg.json", "w"), indent=1)
PY
mkdir -p ckpt
python -m dquartic.cli train --ms2-data-path ms2.npy --ms1-data-path ms1.npy cfg.json 2>&1 | tail -12
ls -la ckpt
python -m dquartic.cli train --ms2-data-path ms2.npy --ms1-data-path ms1.npy --batch-size 4 cfg.json 2>&1 | tail -6
python - <<'PY'
import torch
ck = torch.load("ckpt/dquartic_latest_checkpoint.ckpt", map_location="cpu", weights_only=False)
print("checkpoint keys", sorted(ck.keys()), "epoch", ck["epoch"], "n state entries", len(ck["model_state_dict"]))
assert len(ck["model_state_dict"]) == 396
PY
echo CLI_SMOKE_OK
