"""Host the UNMODIFIED reference model (a checkout of danqu130/RPEFlow) on top of this library.

    from rpeflow_b200 import refhost
    model = refhost.build_rpeflow("/path/to/RPEFlow", device="cuda:0", install=True)   # install() + RPEFlow(cfg).eval()
    out = refhost.forward(model, inputs)        # {'flow_2d': [B,2,H,W], 'flow_3d': [B,3,N]}  (RPEFlow.py:36-99)

``build_rpeflow`` does what eval_withocc.py:30-45 does — read ``conf/test/things.yaml`` into an attribute dict
(omegaconf is not needed: the model only does ``cfgs.a.b`` reads), construct ``models.RPEFlow.RPEFlow`` — and, when
``install=True``, calls ``rpeflow_b200.install.install()`` first so every hot op underneath runs the sm_100a kernels.
The checkout path is an argument: nothing here knows where a reference tree lives.

``synthetic_model_inputs`` makes a FlyingThings3D-shaped batch the way the reference datasets do
(flyingthings3d.py:184-230: perspective point clouds from a depth range, intrinsics (f, cx, cy), two RGB frames,
a [2*bins, H, W] event voxel grid) so that the model can be driven without datasets.
"""
import importlib
import os
import sys
import types

import torch

__all__ = ["AttrDict", "load_config", "import_reference", "build_rpeflow", "synthetic_model_inputs", "forward"]


class AttrDict(dict):
    """dict with attribute reads — stands in for omegaconf.DictConfig (train.py:264-269)."""
    __getattr__ = dict.__getitem__

    @classmethod
    def wrap(cls, x):
        return cls({k: cls.wrap(v) for k, v in x.items()}) if isinstance(x, dict) else x


def load_config(root, rel="conf/test/things.yaml"):
    import yaml
    with open(os.path.join(root, rel)) as f:
        return AttrDict.wrap(yaml.safe_load(f))


_OPTIONAL = ("hdf5plugin", "h5py", "imageio", "skimage", "omegaconf", "matplotlib", "matplotlib.colors", "cv2")


def import_reference(root, stub_optional=False):
    """Put the checkout on sys.path.  stub_optional=True additionally registers empty stand-ins for the data-side
    packages the image does not have (only needed to import event_utils / dsec; models/* needs none of them)."""
    root = os.path.abspath(root)
    if not os.path.isdir(os.path.join(root, "models")):
        raise FileNotFoundError(f"{root} is not an RPEFlow checkout (no models/ directory)")
    if root not in sys.path:
        sys.path.insert(0, root)
    sys.dont_write_bytecode = True                       # a checkout may be read-only
    if stub_optional:
        for name in _OPTIONAL:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules.setdefault(name, types.ModuleType(name))
        if not hasattr(sys.modules["h5py"], "File"):
            sys.modules["h5py"].File = object
        om = sys.modules["omegaconf"]
        if not hasattr(om, "OmegaConf"):
            om.OmegaConf, om.DictConfig = object, dict
        if not hasattr(sys.modules["matplotlib.colors"], "hsv_to_rgb"):
            sys.modules["matplotlib.colors"].hsv_to_rgb = None
    return root


def _patch_mi_for_cpu():
    """models/mutual_info.py:32,84,155,211 draw noise with torch.cuda.FloatTensor, which needs a driver.  On a CPU-only
    box the draw is replaced by randn_like; it only feeds the MI loss, never the flows."""
    mi = importlib.import_module("models.mutual_info")

    def reparametrize(self, mu, logvar):
        std = logvar.mul(0.5).exp_()
        return torch.randn_like(std).mul(std).add_(mu)
    for name in ("Mutual_info_reg_2D", "Mutual_info_reg_2D_Event", "Mutual_info_reg_3D", "Mutual_info_reg_3D_Event"):
        getattr(mi, name).reparametrize = reparametrize


def build_rpeflow(root, device="cuda", install=True, seed=0, config="conf/test/things.yaml", patch_events=False):
    """-> models.RPEFlow.RPEFlow(cfg.model).eval() on `device`, random-init weights under `seed`.
    install=True wires this library in first (rpeflow_b200.install.install); install=False leaves the checkout as
    it is (its own extensions if built, else its torch fallbacks)."""
    root = import_reference(root)
    if install:
        from . import install as _install
        _install.install(patch_python=True, patch_events=patch_events)
    cfg = load_config(root, config).model
    rpeflow = importlib.import_module("models.RPEFlow")
    dev = torch.device(device)
    if dev.type != "cuda":
        _patch_mi_for_cpu()
    torch.manual_seed(seed)
    model = rpeflow.RPEFlow(cfg).to(dev).eval()
    return model


def synthetic_model_inputs(batch, height=540, width=960, n_points=8192, event_bins=10, focal=1050.0, max_depth=35.0,
                           seed=0, first_sample=0, with_voxel=True):
    """One FlyingThings3D-shaped batch on the host.  images [B,6,H,W] in 0..255 (RGB1|RGB2), pcs [B,6,N] (xyz1|xyz2,
    perspective camera coordinates), intrinsics [B,3] = (f, cx, cy) (flyingthings3d.py:184), event_voxel
    [B,2*bins,H,W] (a sparse non-negative count grid, as eventsToVoxel produces)."""
    cx, cy = (width - 1) / 2.0, (height - 1) / 2.0
    f = focal * width / 960.0
    out = {"images": torch.empty(batch, 6, height, width), "pcs": torch.empty(batch, 6, n_points),
           "intrinsics": torch.tensor([[f, cx, cy]] * batch, dtype=torch.float32)}
    if with_voxel:
        out["event_voxel"] = torch.zeros(batch, 2 * event_bins, height, width)
    for i in range(batch):
        g = torch.Generator().manual_seed(7000 + seed * 100003 + first_sample + i)
        # smooth images (a few low-frequency waves + noise) so that the correlation volumes are not pure noise
        yy = torch.linspace(0, 1, height).view(1, height, 1)
        xx = torch.linspace(0, 1, width).view(1, 1, width)
        for half in range(2):
            ph = torch.rand(3, 1, 1, generator=g) * 6.28
            fr = torch.rand(3, 1, 1, generator=g) * 20 + 4
            img = 0.5 + 0.35 * torch.sin(fr * xx + ph + 0.03 * half) * torch.cos(0.7 * fr * yy - ph)
            img = img + 0.05 * torch.randn(3, height, width, generator=g)
            out["images"][i, 3 * half:3 * half + 3] = (img.clamp(0, 1) * 255.0)
            u = torch.rand(n_points, generator=g) * (width - 1)
            v = torch.rand(n_points, generator=g) * (height - 1)
            z = torch.rand(n_points, generator=g) * (max_depth - 2.0) + 2.0
            out["pcs"][i, 3 * half + 0] = (u - cx) * z / f
            out["pcs"][i, 3 * half + 1] = (v - cy) * z / f
            out["pcs"][i, 3 * half + 2] = z
        if with_voxel:
            n_ev = height * width // 2
            flat = out["event_voxel"][i].view(-1)
            idx = torch.randint(0, flat.numel(), (n_ev,), generator=g)
            flat.index_add_(0, idx, torch.rand(n_ev, generator=g))
    return out


@torch.no_grad()
def forward(model, inputs):
    """eval_withocc.py:57: outputs = model.forward(inputs, is_Train=False)."""
    return model.forward(inputs, is_Train=False)
