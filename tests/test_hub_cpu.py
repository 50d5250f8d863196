"""Weight ingestion: the hub cache layout of azula/hub.py and plugins.adm.load_model end to end
from a cached guided-diffusion ``state_dict`` (no network: the file is placed in the cache the way the reference
would have left it)."""

import hashlib
import os
import pytest
import torch
import zipfile

from types import SimpleNamespace

from oracle.gen_golden_cfg import TINY_ADM

from azula_b200 import hub
from azula_b200.plugins import adm


@pytest.fixture
def cache(tmp_path):
    old = hub.get_hub_dir()
    hub.set_hub_dir(str(tmp_path))
    yield tmp_path
    hub.set_hub_dir(old)


def test_cache_layout_matches_reference(cache):
    url = "https://openaipublic.blob.core.windows.net/diffusion/jul-2021/256x256_diffusion_uncond.pt"
    want = os.path.join(str(cache), "https.openaipublic.blob.core.windows.net.diffusion.jul.2021.256x256_diffusion_uncond.pt")
    assert hub.cache_path(url) == want  # azula/hub.py:63-64: runs of non [a-zA-Z0-9_] become one dot


def test_hash_check_and_extract(cache):
    url = "https://example.org/archive.zip"
    path = hub.cache_path(url)
    with zipfile.ZipFile(path, "w") as z:
        z.writestr("inner/checkpoint-last.pth", b"payload")
    digest = hashlib.sha256(open(path, "rb").read()).hexdigest()
    assert hub.download(url, hash_prefix=f"sha256:{digest[:12]}", quiet=True) == path
    with pytest.raises(AssertionError):
        hub.download(url, hash_prefix="sha256:000000", quiet=True)
    xd = hub.download(url, extract=True, quiet=True)
    assert xd == path + "+x" and open(os.path.join(xd, "inner", "checkpoint-last.pth"), "rb").read() == b"payload"
    assert hub.download(url, extract=True, quiet=True) == xd  # second call reuses the directory


def test_load_model_from_cached_checkpoint(cache, monkeypatch):
    torch.manual_seed(0)
    src = adm.make_model(**TINY_ADM)
    adm.seed_parameters(src.backbone, seed=5)
    url = "https://example.org/tiny_adm.pt"
    torch.save(src.backbone.state_dict(), hub.cache_path(url))
    digest = hashlib.sha256(open(hub.cache_path(url), "rb").read()).hexdigest()
    card = SimpleNamespace(url=url, hash=f"sha256:{digest[:16]}", config=TINY_ADM)
    monkeypatch.setattr(adm, "load_cards", lambda _: {"tiny": card})
    den = adm.load_model("tiny")
    assert not den.training
    for (k, a), (_, b) in zip(den.backbone.state_dict().items(), src.backbone.state_dict().items()):
        assert torch.equal(a, b), k
    x = torch.randn(2, 3, 16, 16)
    with torch.no_grad():
        assert torch.equal(den(x, torch.tensor(0.5)).mean, src.eval()(x, torch.tensor(0.5)).mean)


def test_bad_download_is_not_cached_and_drive_links_need_gdown(cache, monkeypatch):
    """a download whose hash does not match must not stay in the cache; Google-Drive links go through
    gdown as in the reference (azula/hub.py:78-79) or fail loudly -- never cache the HTML interstitial."""
    import sys

    url = "https://example.org/weights.pt"
    monkeypatch.setattr(torch.hub, "download_url_to_file", lambda u, dst, progress=True: open(dst, "wb").write(b"garbage"))
    with pytest.raises(AssertionError):
        hub.download(url, hash_prefix="sha256:ffffffff", quiet=True)
    assert not os.path.exists(hub.cache_path(url))
    monkeypatch.setitem(sys.modules, "gdown", None)  # import gdown -> ImportError
    with pytest.raises(RuntimeError, match="gdown"):
        hub.download("https://drive.google.com/uc?id=abc", quiet=True)
    assert not os.path.exists(hub.cache_path("https://drive.google.com/uc?id=abc"))
    fake = SimpleNamespace(download=lambda u, dst, quiet=False: open(dst, "wb").write(b"ok"))
    monkeypatch.setitem(sys.modules, "gdown", fake)
    assert open(hub.download("https://drive.google.com/uc?id=abc", quiet=True), "rb").read() == b"ok"
