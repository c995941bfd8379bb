"""Build-container-only script (needs /root/reference): exports the reference checkpoints to .tdmw
and writes the committed golden fixtures under tests/golden/.  The reference is Python, so it is
imported and run here on CPU; it cannot travel to the GPU box, the fixtures can.

    python oracle/gen_golden.py            # writes tandem_b200/weights/*.tdmw and tests/golden/*.npz

Sources:
  * golden inputs/outputs of the deployed model (abl04, depth_num (48,4,4)):
        tandem/exported/tandem/sample_inputs.pt, tandem/exported/tandem_512x320/sample_inputs.pt
    (written by cva_mvsnet/export_model.py:55-65,159-180; consumed by test_dr_mvsnet, dr_mvsnet.cpp:376-556)
  * reference model run on CPU for the benchmark configuration abl03 (48,32,8), for which nothing is shipped.
"""
import os
import sys
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, REPO)
from tandem_b200.weights_io import save_tdmw  # noqa: E402


def import_reference():
    sys.path.insert(0, f"{REF}/cva_mvsnet")
    pkg = types.ModuleType("models")
    pkg.__path__ = [f"{REF}/cva_mvsnet/models"]
    sys.modules["models"] = pkg  # bypass models/__init__.py (needs pytorch_lightning)
    from models.cva_mvsnet import CvaMVSNet, StageTensor
    return CvaMVSNet, StageTensor


def load_ckpt(name):
    ck = torch.load(f"{REF}/cva_mvsnet/pretrained/ablation/{name}.ckpt", map_location="cpu", weights_only=False)
    sd = {k[len("cva_mvsnet."):]: v for k, v in ck["state_dict"].items()}
    return sd, tuple(ck["hparams"]["MODEL.DEPTH_NUM"]), bool(ck["hparams"]["MODEL.VIEW_AGGREGATION"])


def export(name, out):
    sd, dn, va = load_ckpt(name)
    tensors = {k: v.numpy() for k, v in sd.items() if not k.endswith("num_batches_tracked")}
    save_tdmw(out, tensors, dn, va)
    print(f"wrote {out}: {len(tensors)} tensors, depth_num={dn}, va={va}")


def build_model(name):
    CvaMVSNet, _ = import_reference()
    sd, dn, va = load_ckpt(name)
    net = CvaMVSNet(depth_num=dn, view_aggregation=va).eval()
    net.load_state_dict(sd)
    return net


def read_sample(path):
    t = torch.jit.load(path, map_location="cpu")
    g = lambda n: getattr(t, n)
    return dict(image=g("image"), K=[g(f"intrinsic_matrix.stage{s}") for s in (1, 2, 3)],
                c2w=g("cam_to_world"), dmin=g("depth_min"), dmax=g("depth_max"), disc=g("discard_percentage"),
                out={s: (g(f"outputs.stage{s}.depth"), g(f"outputs.stage{s}.confidence")) for s in (1, 2, 3)})


def run_ref(net, s, order):
    _, StageTensor = import_reference()
    with torch.no_grad():
        return net(s["image"][:, order], StageTensor(*s["K"]), s["c2w"][:, order], s["dmin"], s["dmax"], s["disc"])


def main():
    torch.set_num_threads(os.cpu_count())
    os.makedirs(f"{REPO}/tandem_b200/weights", exist_ok=True)
    os.makedirs(f"{REPO}/tests/golden", exist_ok=True)
    export("abl03_view_aggregation", f"{REPO}/tandem_b200/weights/abl03_view_aggregation.tdmw")
    export("abl04_fewer_depth_planes", f"{REPO}/tandem_b200/weights/abl04_fewer_depth_planes.tdmw")

    net03 = build_model("abl03_view_aggregation")
    net04 = build_model("abl04_fewer_depth_planes")
    for tag, folder in (("640x480", "tandem"), ("512x320", "tandem_512x320")):
        s = read_sample(f"{REF}/tandem/exported/{folder}/sample_inputs.pt")
        V = s["image"].shape[1]
        order = [V - 2] + list(range(V - 2)) + [V - 1]      # export_model.py:126-127 (ref = V-2 first)
        img = s["image"][0]                                  # (V,3,H,W) fp32 RGB/255, window order
        u8 = torch.round(img * 255.0).to(torch.uint8)
        assert torch.equal(u8.float() / 255.0, img), "golden image is exactly u8/255"
        bgr = u8.permute(0, 2, 3, 1).flip(-1).contiguous().numpy()  # (V,H,W,3) BGR: what CallAsync receives
        d = dict(bgr=bgr, ref_index=np.int32(V - 2),
                 K1=s["K"][0][0].numpy(), K2=s["K"][1][0].numpy(), K3=s["K"][2][0].numpy(),
                 c2w=s["c2w"][0].numpy(), depth_min=np.float32(s["dmin"][0]), depth_max=np.float32(s["dmax"][0]),
                 discard=np.float32(s["disc"][0]))
        for st in (1, 2, 3):
            d[f"abl04_stage{st}_depth"] = s["out"][st][0][0].numpy()
            d[f"abl04_stage{st}_confidence"] = s["out"][st][1][0].numpy()
        # self-check: the python reference (abl04) reproduces the shipped goldens
        o4 = run_ref(net04, s, order)
        for st in (1, 2, 3):
            e = (o4[st - 1].depth[0] - s["out"][st][0][0]).abs().mean().item()
            print(f"{tag} abl04 stage{st} ref-model vs shipped golden mean-abs {e:.3e}")
            assert e < 1e-4
        d["abl04_stage3_depth_dense"] = o4[2].depth_dense[0].numpy()
        d["abl04_stage3_confidence_dense"] = o4[2].confidence_dense[0].numpy()
        # benchmark config (48,32,8): reference model output is the ground truth
        o3 = run_ref(net03, s, order)
        for st in (1, 2, 3):
            d[f"abl03_stage{st}_depth"] = o3[st - 1].depth[0].numpy()
            d[f"abl03_stage{st}_confidence"] = o3[st - 1].confidence[0].numpy()
            d[f"abl03_stage{st}_depth_dense"] = o3[st - 1].depth_dense[0].numpy()
            d[f"abl03_stage{st}_confidence_dense"] = o3[st - 1].confidence_dense[0].numpy()
        out = f"{REPO}/tests/golden/sample_{tag}.npz"
        np.savez_compressed(out, **d)
        print(f"wrote {out} ({os.path.getsize(out) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
