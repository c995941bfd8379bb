"""Export a CVA-MVSNet Lightning checkpoint (.ckpt) to the TDMW container read by tandem_b200.
    python tools/export_weights.py <in.ckpt> <out.tdmw>
Needs torch only to unpickle the checkpoint; no reference code is imported."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tandem_b200.weights_io import save_tdmw  # noqa: E402


def main(src, dst):
    ck = torch.load(src, map_location="cpu", weights_only=False)
    sd = {k[len("cva_mvsnet."):]: v.numpy() for k, v in ck["state_dict"].items()
          if k.startswith("cva_mvsnet.") and not k.endswith("num_batches_tracked")}
    save_tdmw(dst, sd, tuple(ck["hparams"]["MODEL.DEPTH_NUM"]), bool(ck["hparams"]["MODEL.VIEW_AGGREGATION"]))
    print(f"wrote {dst}: {len(sd)} tensors")


if __name__ == "__main__":
    main(*sys.argv[1:3])
