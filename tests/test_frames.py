"""Frame front-end: centre crop + Pillow-exact bilinear resize.  Integer work: every check is
bit-exact.  CPU tests pin the oracle (and the library's host-side tables) to the fixture written from the real
torchvision / Pillow transforms; GPU tests compare the kernels with the oracle through the C ABI."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import frame_oracle as FO
from oracle.make_golden_frames import GEOMETRIES, make_frame


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(golden_dir / "frames_golden.npz")


# ------------------------------------------------------------------------------------------------ CPU: oracle pinned
@pytest.mark.parametrize("i", range(len(GEOMETRIES)))
def test_oracle_matches_pillow_fixture(golden, i):
    h, w, hf, wf = GEOMETRIES[i]
    h, w = int(h), int(w)
    assert np.array_equal(golden["geometries"][i], GEOMETRIES[i])
    box = FO.center_crop_box(h, w, hf, wf)
    assert tuple(golden[f"box_{i}"]) == box
    img = make_frame(h, w, 100 + i)
    top, left, ch, cw = box
    res = FO.resize_bilinear_u8(img[top:top + ch, left:left + cw])
    assert res.shape == (224, 224, 3) and res.dtype == np.uint8
    assert np.array_equal(res[:16], golden[f"head_{i}"])
    assert hashlib.sha256(res.tobytes()).digest() == golden[f"sha_{i}"].tobytes()


def test_oracle_matches_installed_pillow():
    """live check against the libraries in this image (skipped where they are missing)"""
    Image = pytest.importorskip("PIL.Image")
    T = pytest.importorskip("torchvision.transforms")
    rng = np.random.default_rng(7)
    for h, w in [(97, 131), (300, 280), (225, 223), (640, 360)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        cc = T.CenterCrop((0.8 * h, 0.8 * w))(Image.fromarray(img))
        ref = np.asarray(T.Resize((224, 224))(cc))
        got = FO.crop_resize_frames(img[None])[0]
        assert np.array_equal(ref, got)


@pytest.mark.parametrize("size", [80, 179, 223, 224, 225, 622, 864, 1536, 3000])
def test_host_tables_match_oracle(size):
    """the library's host-side coefficient tables (C doubles) equal the oracle's (numpy doubles), tap for tap"""
    from sais_b200 import frames as F
    bounds, kk, ksize = FO.resample_coeffs(size, 224)
    tab = F.resize_table(size)
    assert tab.size == 448 + ksize * 224
    assert np.array_equal(tab[:448].reshape(224, 2), bounds)
    assert np.array_equal(tab[448:].reshape(ksize, 224).T, kk)
    assert int(kk.sum(1).min()) >= (1 << 22) - ksize and int(kk.sum(1).max()) <= (1 << 22) + ksize  # weights sum to ~1


def test_crop_box_host_matches_oracle():
    from sais_b200 import frames as F
    for h in (100, 224, 281, 355, 480, 721, 1080):
        for w in (130, 224, 333, 501, 854, 1285, 1920):
            for hf, wf in ((0.8, 0.8), (0.8, 0.7)):
                assert F.center_crop_box(h, w, hf, wf) == FO.center_crop_box(h, w, hf, wf)
    assert F.get_crop_dims("VUA_Gronau") == (0.8, 0.7) and F.get_crop_dims("VUA") == (0.8, 0.8)


def test_frames_refuse_cpu_tensors():
    from sais_b200 import SaisError, frames as F
    with pytest.raises(SaisError):
        F.crop_resize(torch.zeros(1, 300, 300, 3, dtype=torch.uint8))


# ------------------------------------------------------------------------------------------------ GPU: kernels vs oracle
@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(GEOMETRIES)))
def test_gpu_crop_resize_bit_exact(dev, golden, i):
    from sais_b200 import frames as F
    h, w, hf, wf = GEOMETRIES[i]
    h, w = int(h), int(w)
    n = 3 if h * w < 600 * 900 else 2
    imgs = np.stack([make_frame(h, w, 100 + i + 50 * j) for j in range(n)])
    got = F.crop_resize(torch.from_numpy(imgs).to(dev), hf, wf).cpu().numpy()
    ref = FO.crop_resize_frames(imgs, hf, wf)
    assert np.array_equal(got, ref)
    assert hashlib.sha256(got[0].tobytes()).digest() == golden[f"sha_{i}"].tobytes()  # = real Pillow output


@pytest.mark.gpu
def test_gpu_crop_resize_misaligned_views_and_edges(dev):
    """frame sizes whose byte counts are not multiples of 16, batch slices that start mid-buffer, a crop that touches the
    first / last byte of the buffer, upscaling, an explicit box, and N = 0"""
    from sais_b200 import frames as F
    rng = np.random.default_rng(3)
    imgs = rng.integers(0, 256, (5, 281, 501, 3), dtype=np.uint8)
    t = torch.from_numpy(imgs).to(dev)
    ref = FO.crop_resize_frames(imgs)
    assert np.array_equal(F.crop_resize(t[1:4]).cpu().numpy(), ref[1:4])
    full = F.crop_resize(t, box=(0, 0, 281, 501)).cpu().numpy()  # whole frame: first and last chunk hit the buffer ends
    assert np.array_equal(full, np.stack([FO.resize_bilinear_u8(f) for f in imgs]))
    small = rng.integers(0, 256, (2, 60, 75, 3), dtype=np.uint8)  # 48 x 60 crop -> upscaled
    assert np.array_equal(F.crop_resize(torch.from_numpy(small).to(dev)).cpu().numpy(), FO.crop_resize_frames(small))
    assert F.crop_resize(t[:0]).shape == (0, 224, 224, 3)
    with pytest.raises(Exception):
        F.crop_resize(t, box=(0, 0, 300, 501))


@pytest.mark.gpu
def test_gpu_crop_resize_full_size_properties(dev):
    """1080p batch (BASELINE-sized frames): constants stay constant, a 280 x 280 frame's 224 x 224 crop passes through
    unchanged, result is invariant to where in a batch a frame sits, and channels do not mix"""
    from sais_b200 import frames as F
    const = torch.full((2, 1080, 1920, 3), 137, dtype=torch.uint8, device=dev)
    assert bool((F.crop_resize(const) == 137).all())
    ident = torch.randint(0, 256, (3, 280, 280, 3), dtype=torch.uint8, device=dev)
    assert torch.equal(F.crop_resize(ident), ident[:, 28:252, 28:252].contiguous())
    big = torch.randint(0, 256, (6, 1080, 1920, 3), dtype=torch.uint8, device=dev)
    all6 = F.crop_resize(big)
    assert torch.equal(F.crop_resize(big[4:5].contiguous()), all6[4:5])
    red = big.clone()
    red[..., 1:] = 0
    r = F.crop_resize(red)
    assert torch.equal(r[..., 0], all6[..., 0]) and bool((r[..., 1:] == 0).all())
    ref0 = FO.crop_resize_frames(big[:1].cpu().numpy())
    assert np.array_equal(all6[:1].cpu().numpy(), ref0)


@pytest.mark.gpu
def test_gpu_frames_to_embeddings(dev):
    """front-end output feeds forward_u8 directly: same embeddings as handing the ViT the oracle-resized frames"""
    from oracle import sais_oracle as O
    from sais_b200 import frames as F, vision_transformer as vits
    imgs = np.stack([make_frame(360, 640, 900 + j) for j in range(4)])
    model = vits.vit_small(16).to(dev).eval()
    model.load_state_dict(O.make_vit_weights(0))
    a = model.forward_u8(F.crop_resize(torch.from_numpy(imgs).to(dev)))
    b = model.forward_u8(torch.from_numpy(FO.crop_resize_frames(imgs)).to(dev))
    assert torch.equal(a, b)
