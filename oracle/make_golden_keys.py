"""TEST INFRASTRUCTURE ONLY — tests/golden/fullmodel_state_keys.json: key names and shapes of the EXECUTED reference
``prepare_model.fullModel(...).state_dict()`` (main.sh:27 configuration), written in the build container.  The CPU test
``test_loadmodel_round_trip`` compares the product's key set against it (minus ``encoder.*``, the unused timm ViT-B that
the stub replaces by ``nn.Identity`` here and that the product drops on load)."""
from __future__ import annotations

import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
GOLD = ROOT / "tests" / "golden"


def main():
    sys.dont_write_bytecode = True
    from oracle import ref_import as R

    pm = R.load_prepare_model()
    m = R.build_full_model(pm, "RGB-Flow", nclasses=2)
    sd = m.state_dict()
    out, pos = {}, {}
    for k, v in sd.items():  # the 2 x 2000 ParameterDict rows are stored as "<prefix>.*": [count, shape]
        head, _, tail = k.rpartition(".")
        if head in ("frame_pos_embeddings", "clip_pos_embeddings") and tail.isdigit():
            cnt, shape, ids = pos.setdefault(head, [0, list(v.shape), set()])
            assert shape == list(v.shape)
            pos[head][0] += 1
            ids.add(int(tail))
        else:
            out[k] = list(v.shape)
    for head, (cnt, shape, ids) in pos.items():
        assert ids == set(range(cnt))
        out[head + ".*"] = [cnt, shape]
    (GOLD / "fullmodel_state_keys.json").write_text(json.dumps(out, indent=0, sort_keys=True) + "\n")
    print("wrote", GOLD / "fullmodel_state_keys.json", len(out), "keys",
          sum(int(v.numel()) for v in sd.values()), "elements")


if __name__ == "__main__":
    main()
