"""world_size-2 gloo test of the read-sharded path's host logic (no GPU): parameter broadcast, read sharding,
count reduction; the per-rank results put back in read order equal the single-process result."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bonito_b200.crf.model import Model
    from bonito_b200.distributed import broadcast_parameters, gather_counts, shard_reads
    from oracle import synth
    spec = synth.model_spec("fast", n_lstm=1)
    model = Model(synth.model_config(spec))
    if rank == 0:
        model.load_state_dict(synth.state_dict_from_weights(spec, synth.make_weights(spec, seed=5)))
    else:
        for p in model.parameters():
            p.data.zero_()
    broadcast_parameters(model)
    reads = [(f"read{i}", synth.squiggle(1, 600, seed=i)) for i in range(7)]
    mine = list(shard_reads(reads))
    model.eval()
    with torch.inference_mode():
        res = {rid: model(x).sum().item() for rid, x in mine}
    total = gather_counts(sum(x.shape[-1] for _, x in mine))
    torch.save({"res": res, "total": total, "w": model.state_dict()["encoder.4.rnn.weight_hh_l0"].clone()},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path):
    world, port = 2, 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]
    assert torch.equal(outs[0]["w"], outs[1]["w"]) and outs[0]["w"].abs().sum() > 0     # broadcast happened
    assert outs[0]["total"] == outs[1]["total"] == 7 * 600
    assert set(outs[0]["res"]) == {"read0", "read2", "read4", "read6"}
    assert set(outs[1]["res"]) == {"read1", "read3", "read5"}
    # single-process reference
    from bonito_b200.crf.model import Model
    from oracle import synth
    spec = synth.model_spec("fast", n_lstm=1)
    model = Model(synth.model_config(spec))
    model.load_state_dict(synth.state_dict_from_weights(spec, synth.make_weights(spec, seed=5)))
    model.eval()
    merged = {**outs[0]["res"], **outs[1]["res"]}
    with torch.inference_mode():
        for i in range(7):
            want = model(synth.squiggle(1, 600, seed=i)).sum().item()
            assert abs(merged[f"read{i}"] - want) < 1e-3
