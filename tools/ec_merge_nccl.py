#!/usr/bin/env python
"""torchrun check (N GPUs): Extract+Count on every rank's record-aligned shards through libf2q, tables merged by
2fast2q_b200.multi.merge_ec_tables over NCCL, compared with the oracle on the whole stream."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from oracle import oracle as O, synth
f2q = importlib.import_module("2fast2q_b200"); lib = f2q._lib
multi = importlib.import_module("2fast2q_b200.multi"); host = importlib.import_module("2fast2q_b200.fast2q")
rank, lr, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
params = dict(mode="EC", upstream="GTTCAGAGTTCT", downstream="CTGAATAGGCCA", miss_search_up=1, miss_search_down=1)
whole = synth.barseq_reads(4, 30000)
blocks = [whole[o:o + 100_001] for o in range(0, len(whole), 100_001)]
mine = {}
with lib.Engine(lib.make_config(**params), lr) as e:
    for shard, final in multi.rank_shards(host.record_aligned_shards(blocks, 400_000), rank, world):
        e.run(shard)
        for k, v in e.ec_items().items():
            mine[k] = mine.get(k, 0) + v
keys = list(mine)
kb = np.frombuffer(b"".join(keys), dtype=np.uint8) if keys else np.zeros(0, dtype=np.uint8)
ko = np.cumsum([0] + [len(k) for k in keys]).astype(np.uint64)
cn = np.array([mine[k] for k in keys], dtype=np.uint64)
mk, mc = multi.merge_ec_tables(kb, ko, cn)
want, _ = O.extract_count(O.make_config(**params), np.frombuffer(whole, dtype=np.uint8))
ok = dict(zip(mk, mc)) == want
print(f"rank {rank}: local keys {len(mine)}, merged keys {len(mk)}, equals oracle: {ok}", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
