"""Install the B200 quantizers into the reference's namespaces.

src/vqvae.py:8 does `from src.embed import L2Embedding, SeperateEmbedding` and src/tts.py:5 does
`from src.embed import L2Embedding as Embedding`, i.e. the names are bound at import time, so the
replacement assigns the new classes in `src.embed` *and* in any already-imported consumer module.
Nothing in the reference tree is modified.
"""
import sys


_ORIGINALS = {}          # (module name, attribute) -> the reference's own object, kept for uninstall_from_reference()


def _swap(obj, owner_key, attr, new):
    _ORIGINALS.setdefault((owner_key, attr), (obj, getattr(obj, attr)))
    setattr(obj, attr, new)


def install_into_reference(mean_forward=True):
    from .embed import L2Embedding, SeperateEmbedding
    import src.embed as ref_embed                       # the reference package must be importable
    _swap(ref_embed, "src.embed", "L2Embedding", L2Embedding)
    _swap(ref_embed, "src.embed", "SeperateEmbedding", SeperateEmbedding)
    vq = sys.modules.get("src.vqvae")
    if vq is not None:
        _swap(vq, "src.vqvae", "L2Embedding", L2Embedding)
        _swap(vq, "src.vqvae", "SeperateEmbedding", SeperateEmbedding)
        if mean_forward:
            # the run-length collapse that follows the quantizer (src/vqvae.py:218-257) moves to the GPU as well
            from .segment import vqvae_mean_forward
            _swap(vq.VQVAE, "src.vqvae.VQVAE", "mean_forward", vqvae_mean_forward)
    tts = sys.modules.get("src.tts")
    if tts is not None:
        _swap(tts, "src.tts", "Embedding", L2Embedding)
    return L2Embedding, SeperateEmbedding


def uninstall_from_reference():
    """Put back everything install_into_reference() replaced (the reference's own classes and VQVAE.mean_forward)."""
    while _ORIGINALS:
        (_, attr), (obj, orig) = _ORIGINALS.popitem()
        setattr(obj, attr, orig)
