"""Golden vectors for collate_fn (TEST INFRASTRUCTURE ONLY): runs the reference's own src/datasets/collate.py:4-46
(it only needs torch) on seeded synthetic items and stores keys / shapes / checksums in tests/golden/collate.json.

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_collate_golden.py
"""
import importlib.util
import json
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def items(seed, n, T, with_gt, with_video, with_emb, extra):
    """Seeded synthetic dataset items with the reference's item keys (tests/test_data_pipeline.py imports this)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n):
        d = {"mix": torch.randn(1, T, generator=g), "audio_path": f"/data/mix/{seed}_{i}.wav",
             "s1": torch.randn(1, T, generator=g) if with_gt else None,
             "s2": torch.randn(1, T, generator=g) if with_gt else None,
             "s1_video": torch.randn(1, 5, 8, 8, generator=g) if with_video else None,
             "s2_video": torch.randn(1, 5, 8, 8, generator=g) if with_video else None,
             "s1_embedding": torch.randn(1, 16, 5, generator=g) if with_emb else None,
             "s2_embedding": torch.randn(1, 16, 5, generator=g) if with_emb else None}
        if extra:
            d["mix_spectrogram"] = torch.randn(1, 4, 9, generator=g)
            d["not_a_batch_key"] = 123
        out.append(d)
    return out


CASES = [(1, 3, 50, True, False, True, False), (2, 1, 20, False, False, False, False), (3, 4, 33, True, True, True, True)]

if __name__ == "__main__":
    spec = importlib.util.spec_from_file_location("ref_collate", "/root/reference/src/datasets/collate.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cases = []
    for args in CASES:
        batch = mod.collate_fn(items(*args))
        desc = {}
        for k, v in batch.items():
            if v is None:
                desc[k] = None
            elif torch.is_tensor(v):
                desc[k] = {"shape": list(v.shape), "sum": float(v.double().sum())}
            else:
                desc[k] = list(v)
        cases.append({"args": list(args), "batch": desc, "key_order": list(batch.keys())})
    out = os.path.join(HERE, "..", "tests", "golden", "collate.json")
    json.dump({"source": "src/datasets/collate.py (reference)", "cases": cases}, open(out, "w"))
    print("wrote", out)
