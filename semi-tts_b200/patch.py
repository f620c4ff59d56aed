"""Install the B200 quantizers into the reference's namespaces.

src/vqvae.py:8 does `from src.embed import L2Embedding, SeperateEmbedding` and src/tts.py:5 does
`from src.embed import L2Embedding as Embedding`, i.e. the names are bound at import time, so the
replacement assigns the new classes in `src.embed` *and* in any already-imported consumer module.
Nothing in the reference tree is modified.
"""
import sys


def install_into_reference(mean_forward=True):
    from .embed import L2Embedding, SeperateEmbedding
    import src.embed as ref_embed                       # the reference package must be importable
    ref_embed.L2Embedding = L2Embedding
    ref_embed.SeperateEmbedding = SeperateEmbedding
    vq = sys.modules.get("src.vqvae")
    if vq is not None:
        vq.L2Embedding = L2Embedding
        vq.SeperateEmbedding = SeperateEmbedding
        if mean_forward:
            # the run-length collapse that follows the quantizer (src/vqvae.py:218-257) moves to the GPU as well
            from .segment import vqvae_mean_forward
            vq.VQVAE.mean_forward = vqvae_mean_forward
    tts = sys.modules.get("src.tts")
    if tts is not None:
        tts.Embedding = L2Embedding
    return L2Embedding, SeperateEmbedding
