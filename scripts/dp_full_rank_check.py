"""2-GPU check of the vocab-sharded full-catalogue ranking (run under torchrun --nproc-per-node 2): every rank ranks its
OWN sequences; the sharded path (all-gather rows, count per vocabulary shard, all-reduce counts over NCCL) must equal the
unsharded count computed locally."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from tests.helpers import make_batch
from bert4rec_b200.models import BERT4RecModel
from bert4rec_b200.models.components import networks

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
kw = dict(vocab_size=30011, hidden_size=64, num_layers=2, num_attention_heads=2, max_sequence_length=50, inner_dim=64)
B, S, P = 64, 50, 8
model = BERT4RecModel(networks.Bert4RecEncoder(**kw, device=f"cuda:{local}", seed=0))   # same weights on every rank
batch = make_batch(B, S, P, kw["vocab_size"], seed=100 + rank, eval_mode=True)             # different sequences per rank
gt = batch["masked_lm_ids"][:, 0].clone()
sharded = model.full_catalogue_ranks(batch, gt).cpu()
# unsharded reference on this rank's rows
sess = model.store.session(B, S, P)
n = B
t = sess.mlm_hidden()[:n].clone()
_, score, _ = sess.rank_candidates(gt.cuda().view(-1, 1), None, want_ranking=False, want_scores=True)
whole = torch.zeros(n, dtype=torch.int32, device=f"cuda:{local}")
sess.rank_full_ext(t, gt.to(torch.int32).cuda(), score.view(-1).contiguous(), torch.tensor([n, n], dtype=torch.int32, device=f"cuda:{local}"),
                   0, kw["vocab_size"], whole)
ok = torch.equal(whole.cpu() + 1, sharded.to(torch.int32))
flag = torch.tensor([int(ok)], device=f"cuda:{local}")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"vocab-sharded full-catalogue ranking over {world} ranks (NCCL): {'EQUAL to unsharded' if int(flag) else 'MISMATCH'}; "
          f"rank 0 sample ranks {sharded[:8].tolist()}")
dist.destroy_process_group()
sys.exit(0 if int(flag) else 1)
