"""Generates the golden fixtures by running the UNMODIFIED reference classes (imported from /root/reference)
on weights / inputs produced by the oracle's deterministic generators.  Run in the build container only:

    python tests/golden/gen_golden.py [resnet] [vis] [kmeans]

The fixtures travel with the repo; /root/reference does not exist on the GPU box.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")


def gen_resnet():
    from oracle import resnet50_oracle as O
    from src.resnet import resnet50  # the reference
    torch.manual_seed(0)
    sd = O.make_state_dict(0)
    ref = resnet50(pretrained=False).eval()
    ref.load_state_dict(sd, strict=True)
    patches = O.make_patches(7, 2)
    with torch.no_grad():
        x = O.preprocess(patches)
        feat = ref.forward_extract(x)
        # the same through the reference's own preprocessing objects (torchvision transforms)
        from torchvision import transforms
        tv = torch.nn.Sequential(transforms.ConvertImageDtype(torch.float),
                                 transforms.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225]))
        x_tv = torch.stack([tv(p.permute(2, 0, 1)) for p in patches])
        assert torch.equal(x_tv, x), "preprocessing restatement differs from torchvision transforms"
        feat64 = ref.double().forward_extract(x.double())
    np.savez_compressed(os.path.join(HERE, "resnet50_golden.npz"), weights_seed=0, patches_seed=7, n=2,
                        features=feat.numpy(), features_fp64=feat64.numpy())
    print("resnet golden:", feat.shape, float(feat.abs().mean()))


if __name__ == "__main__":
    what = sys.argv[1:] or ["resnet", "vis", "kmeans"]
    for w in what:
        globals()["gen_" + w]()
