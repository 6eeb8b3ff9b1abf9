"""Host logic of the synthetic dataset that stands in for data_c.LibriPartyDataset (data_c.py:60-83,156-175): item dict,
shapes, the flag=False / flag=True piece protocol, determinism."""
import torch

from speech_enhancement_mi_b200.data_synth import SyntheticPartyDataset


def test_item_dict_and_piece_protocol():
    d = SyntheticPartyDataset(size=6, utterance_seconds=5.0, max_length=60000, seed=0)
    d.init_seed(3)
    first_of_mixture = True
    covered = 0
    for i in range(len(d)):
        it = d[i]
        assert set(it) == {"mix", "source", "noise", "length", "flag"}
        L = int(it["length"])
        assert 16000 <= L < 60000
        assert it["mix"].shape == (3, L) and it["noise"].shape == (3, L) and it["source"].shape == (1, 1, L)
        assert it["mix"].dtype == torch.float32 and float(it["mix"].abs().max()) <= 0.95 + 1e-6
        # train.py:186-188: source.squeeze(1)[:, 0] is the clean mic-0 signal; mix = source + noise on mic 0
        assert torch.allclose(it["mix"][0], it["source"][0, 0] + it["noise"][0], atol=1e-6)
        assert it["flag"] == (not first_of_mixture)  # a mixture's first piece resets the model, later pieces continue it
        covered += L
        first_of_mixture = not d.buffer  # buffer empty -> the next item starts a new mixture
        if first_of_mixture:
            assert covered <= 80000
            covered = 0


def test_seed_makes_the_stream_of_items_reproducible():
    a, b = SyntheticPartyDataset(size=3), SyntheticPartyDataset(size=3)
    a.init_seed(5)
    b.init_seed(5)
    for i in range(3):
        x, y = a[i], b[i]
        assert torch.equal(x["mix"], y["mix"]) and x["flag"] == y["flag"]
    b.init_seed(6)
    b.set_attribute("train")  # train.py:167-170: new seed, buffer emptied
    c = SyntheticPartyDataset(size=3)
    c.init_seed(5)
    assert not torch.equal(c[0]["mix"][:, :16000], b[0]["mix"][:, :16000])
