"""Golden draw sequences of the reference's MultiTraversalBalancedSampler (mtgs/dataset/utils/sampler.py:27-58), produced
by the reference itself (the module only needs ``random`` and numpy; it is loaded by file path).  Run in the build
container: python tests/golden/make_sampler_golden.py -> sampler_reference_golden.json"""
import importlib.util
import json
import os
import random
import types

REF = "/root/reference/mtgs/dataset/utils/sampler.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sampler_reference_golden.json")


def main():
    spec = importlib.util.spec_from_file_location("ref_sampler", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    cases = {}
    for name, travel_ids, seed, draws in (("three_traversals", [0] * 7 + [1] * 4 + [7] * 9, 123, 60),
                                          ("interleaved", [3, 1, 3, 2, 1, 2, 3, 3, 1, 2, 2, 2], 7, 50),
                                          ("single", [5] * 6, 1, 20)):
        ds = types.SimpleNamespace(_dataparser_outputs=types.SimpleNamespace(travel_ids=travel_ids))
        random.seed(seed)
        s = ref.MultiTraversalBalancedSampler(ds)
        cases[name] = {"travel_ids": travel_ids, "seed": seed, "sequence": [int(s.get_next_image_idx()) for _ in range(draws)]}
    json.dump(cases, open(OUT, "w"), indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
