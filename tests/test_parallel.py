"""Tensor parallelism: the sharding plan and the exchange arithmetic, checked on CPU with world_size-2/4 `gloo`
process groups (the N>1 host path), and on real GPUs with NCCL when two or more are visible.

The CPU test runs a complete sharded forward of the tiny model with the ORACLE as each rank's compute: every rank
slices its weights with the product's `parallel.shard_of/slice_canonical`, exchanges unrounded f64 partials through
`all_reduce`, and must reproduce the unsharded oracle's logits bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _tp_oracle_worker(rank, world, port, path, tokens, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["OMP_NUM_THREADS"] = "2"
    import torch
    import torch.distributed as dist
    torch.set_num_threads(1)
    from ggufb200 import gguf_reader as G, parallel
    from ggufb200.model import HParams
    from oracle import oracle as O
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    f = G.GGUFFile(path)
    hp = HParams.from_gguf(f)
    parallel.check_divisible(hp, world)
    nh, nkv, hd = hp.n_head // world, hp.n_kv // world, hp.head_dim

    def W(name):
        t = f.tensors[name]
        raw, rows, k = parallel.slice_canonical(np.asarray(f.data(name)), t.ggml_type, t.ne[0], t.ne[1], parallel.shard_of(name, hp, world, rank))
        return t.ggml_type, raw, rows, k

    def norm(name):
        t = f.tensors[name]
        return O.dequantize(f.data(name), t.ggml_type, t.n_elements)

    def mv(name, x):
        qt, raw, rows, k = W(name)
        return O.matmul(qt, raw, rows, k, x, mode="canon")

    def mv_reduced(name, x):                       # row-split projection: f64 partials, all-reduce, ONE rounding
        qt, raw, rows, k = W(name)
        part = torch.from_numpy(O.matvec_f64(qt, raw, rows, k, x))
        dist.all_reduce(part)
        return part.numpy().astype(np.float32)

    n_ctx = 32
    kc = np.zeros((hp.n_layer, n_ctx, nkv * hd), np.uint16)
    vc = np.zeros_like(kc)
    emb = f.tensors["token_embd.weight"]
    rb = G.row_bytes(emb.ggml_type, hp.d)
    logits = None
    for pos, tok in enumerate(tokens):
        x = O.dequantize(f.data("token_embd.weight")[tok * rb:(tok + 1) * rb], emb.ggml_type, hp.d)
        for l in range(hp.n_layer):
            p = f"blk.{l}."
            h = O.rms_norm(x, norm(p + "attn_norm.weight"), hp.eps)
            tab = O.rope_table_canon(pos, hp.n_rot, hp.rope_base)
            q = O.rope_apply(mv(p + "attn_q.weight", h), nh, hd, hp.n_rot, tab)
            k = O.rope_apply(mv(p + "attn_k.weight", h), nkv, hd, hp.n_rot, tab)
            kc[l, pos] = O.fp32_to_fp16(k)
            vc[l, pos] = O.fp32_to_fp16(mv(p + "attn_v.weight", h))
            a = O.attn_decode(q, kc[l], vc[l], nh, nkv, hd, pos + 1, mode="canon")
            x = x + mv_reduced(p + "attn_output.weight", a)
            h = O.rms_norm(x, norm(p + "ffn_norm.weight"), hp.eps)
            act = O.swiglu(mv(p + "ffn_gate.weight", h), mv(p + "ffn_up.weight", h), mode="canon")
            x = x + mv_reduced(p + "ffn_down.weight", act)
        logits = mv("output.weight", O.rms_norm(x, norm("output_norm.weight"), hp.eps))
    # vocabulary-sharded arg-max through the same sortable key the CUDA path uses
    v = hp.vocab // world
    li = int(np.argmax(logits))
    b = np.float32(logits[li]).view(np.uint32)
    mono = (~b & 0xFFFFFFFF) if b & 0x80000000 else (b | 0x80000000)
    key = ((int(mono) << 32) | (0xFFFFFFFF - (li + rank * v))) ^ 0x8000000000000000
    key = key - (1 << 64) if key >= (1 << 63) else key
    kt = torch.tensor([key], dtype=torch.int64)
    dist.all_reduce(kt, op=dist.ReduceOp.MAX)
    k2 = (int(kt[0]) + (1 << 64)) % (1 << 64) ^ 0x8000000000000000
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), logits=logits, tok=0xFFFFFFFF - (k2 & 0xFFFFFFFF))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_forward_with_gloo_equals_unsharded_oracle(oracle, model_dir, tmp_path, world):
    import torch.multiprocessing as mp
    from ggufb200 import synth
    from dataclasses import replace
    # 16 heads x 64, 4 KV heads, ff 3072, vocab 8192: every K-slice stays a multiple of 256 for tp = 2 and 4
    cfg = replace(synth.PRESETS["medium"], n_layer=2, ff=3072)
    path = os.path.join(model_dir, "tp-med2.gguf")
    if not os.path.exists(path):
        synth.write_gguf(path, cfg, "Q4_K_M", seed=0xB200)
    tokens = [1, 300, 301]
    mp.spawn(_tp_oracle_worker, args=(world, _free_port(), path, tokens, str(tmp_path)), nprocs=world, join=True)
    ref = oracle.OracleLlama(path, n_ctx=32, mode="canon")
    for i, t in enumerate(tokens):
        full = ref.forward(t, i)
    got = np.concatenate([np.load(tmp_path / f"rank{r}.npz")["logits"] for r in range(world)])
    assert np.array_equal(got.view(np.uint32), full.view(np.uint32))
    assert all(int(np.load(tmp_path / f"rank{r}.npz")["tok"]) == int(np.argmax(full)) for r in range(world))


def test_shard_plan_covers_every_weight_once_and_rejects_bad_sizes():
    from ggufb200 import parallel, synth
    from ggufb200.model import HParams
    hp = HParams(n_layer=80, d=8192, ff=28672, n_head=64, n_kv=8, head_dim=128, n_rot=128, eps=1e-5, rope_base=5e5, vocab=128256, ctx_train=8192)
    for tp in (1, 2, 4, 8):
        parallel.check_divisible(hp, tp)
        for name, full in (("blk.3.attn_q.weight", 8192), ("blk.3.attn_k.weight", 1024), ("blk.3.attn_output.weight", 8192),
                           ("blk.3.ffn_up.weight", 28672), ("blk.3.ffn_down.weight", 28672), ("output.weight", 128256)):
            if tp == 1:
                assert parallel.shard_of(name, hp, tp, 0).kind == parallel.FULL
                continue
            spans = [parallel.shard_of(name, hp, tp, r) for r in range(tp)]
            assert spans[0].lo == 0 and spans[-1].hi == full and all(a.hi == b.lo for a, b in zip(spans, spans[1:]))
            if spans[0].kind == parallel.COLS:
                assert all(s.lo % 256 == 0 for s in spans)     # K-quant super-block boundary
        assert parallel.shard_of("blk.0.attn_norm.weight", hp, tp, 0).kind == parallel.FULL
        assert parallel.shard_of("token_embd.weight", hp, tp, 0).kind == parallel.FULL
    with pytest.raises(ValueError):
        parallel.check_divisible(hp, 16)                        # 8 KV heads cannot be split 16 ways
    hp8 = HParams(32, 4096, 14336, 32, 8, 128, 128, 1e-5, 5e5, 128256, 8192)
    parallel.check_divisible(hp8, 8)


def _tp_gpu_worker(rank, world, port, path, n_new, out_dir, exchange):
    sys.path.insert(0, ROOT)
    os.environ["GGB_TP_EXCHANGE"] = exchange
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from ggufb200.model import Engine
    eng = Engine(path, n_ctx=128, device=rank, tp_rank=rank, tp_size=world)
    assert (eng.peer is not None) == (exchange == "peer")
    eng.warmup()
    eng.reset()
    eng.prefill([1, 300, 301, 302, 303])
    toks, logits = [], []
    for i in range(n_new):
        logits.append(eng.last_logits())
        toks.append(eng.tokens(i + 1)[i])
        if i + 1 < n_new:
            eng.decode(1)
    np.savez(os.path.join(out_dir, f"gpu{rank}.npz"), toks=np.array(toks), logits=np.stack(logits))
    eng.close()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_tensor_parallel_engine_bit_exact_vs_oracle(oracle, model_dir, tmp_path, exchange):
    """exchange = peer: the all-reduce fused into the row-split GEMVs over NVLink peer memory (csrc/peer.cu);
    nccl: the library collective.  Both must reproduce the single-process canon oracle bit for bit."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    from ggufb200 import synth
    from dataclasses import replace
    world = 2
    cfg = replace(synth.PRESETS["medium"], n_layer=4)            # heads 16, kv 4, ff 2816 -> 1408 = 5.5 x 256: adjust
    cfg = replace(cfg, ff=3072)
    path = os.path.join(model_dir, "tp-medium.gguf")
    if not os.path.exists(path):
        synth.write_gguf(path, cfg, "Q4_K_M", seed=0xB200)
    n_new = 24
    mp.spawn(_tp_gpu_worker, args=(world, _free_port(), path, n_new, str(tmp_path), exchange), nprocs=world, join=True)
    ref = oracle.OracleLlama(path, n_ctx=128, mode="canon")
    ref_toks, ref_logits = ref.greedy([1, 300, 301, 302, 303], n_new, return_logits=True)
    parts = [np.load(tmp_path / f"gpu{r}.npz") for r in range(world)]
    assert list(parts[0]["toks"]) == list(parts[1]["toks"]) == ref_toks
    got = np.concatenate([p["logits"] for p in parts], axis=1)
    assert np.array_equal(got.view(np.uint32), np.stack(ref_logits).view(np.uint32))


TP_PROMPTS = [[1, 300, 301, 302, 303], [1, 310, 311], [1, 320, 321, 322, 323, 324, 325, 326, 327]]


def _tp_batch_worker(rank, world, port, path, n_new, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from ggufb200.model import Engine
    eng = Engine(path, n_ctx=128, device=rank, tp_rank=rank, tp_size=world, n_slots=len(TP_PROMPTS))
    assert eng.batch_capable
    eng.warmup()
    for s, p in enumerate(TP_PROMPTS):
        eng.slots[s].reset()
        eng.slots[s].prefill(p)               # >= 4 tokens: 16-per-pass batched prompt path, f64 all-reduce per row-split projection
    last = [eng.slots[s].read_last_token() for s in range(len(TP_PROMPTS))]
    got = [[t] for t in last]
    bd = eng.batch
    h = bd.launch([(s, last[s], eng.slots[s].n_past) for s in range(len(TP_PROMPTS))])
    for _ in range(n_new - 2):
        h2 = bd.launch_chained()
        for s, t in enumerate(bd.collect(h)):
            got[s].append(t)
        h = h2
    for s, t in enumerate(bd.collect(h)):
        got[s].append(t)
    rows = torch.stack([bd.logits_row_tensor(b) for b in range(len(TP_PROMPTS))]).cpu().numpy()
    np.savez(os.path.join(out_dir, f"tpb{rank}.npz"), toks=np.array(got), logits=rows)
    eng.close()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_tensor_parallel_batched_decode_bit_exact_vs_oracle(oracle, model_dir, tmp_path):
    """several sequences advance together on a tensor-parallel engine (BatchDecoder under tp: f64 partials of the row-split
    projections all-reduced before the one rounding, arg-max as one sortable key per token), one step ahead of the host:
    every sequence's tokens and final logits are the single-process canon oracle's, bit for bit"""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    from ggufb200 import synth
    from dataclasses import replace
    world = 2
    cfg = replace(synth.PRESETS["medium"], n_layer=4, ff=3072)
    path = os.path.join(model_dir, "tp-medium.gguf")
    if not os.path.exists(path):
        synth.write_gguf(path, cfg, "Q4_K_M", seed=0xB200)
    n_new = 12
    mp.spawn(_tp_batch_worker, args=(world, _free_port(), path, n_new, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"tpb{r}.npz") for r in range(world)]
    assert np.array_equal(parts[0]["toks"], parts[1]["toks"])
    logits = np.concatenate([p["logits"] for p in parts], axis=1)
    for s, prompt in enumerate(TP_PROMPTS):
        ref = oracle.OracleLlama(path, n_ctx=128, mode="canon")
        ref_toks, ref_logits = ref.greedy(prompt, n_new, return_logits=True)
        assert list(parts[0]["toks"][s]) == ref_toks, f"sequence {s}"
        assert np.array_equal(logits[s].view(np.uint32), np.asarray(ref_logits[-1]).view(np.uint32)), f"sequence {s}"


def _tp_serve_worker(rank, world, port, path, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    from fake_engine import OracleEngine
    from ggufb200.tp_serve import TPLeader, follower_loop
    eng = OracleEngine(path, n_ctx=64, n_slots=2)
    if rank == 0:
        lead = TPLeader(eng, dist)
        s1, s0 = lead.slots[1], lead.slots[0]
        s1.reset(); s1.prefill([1, 300, 301, 302]); s1.decode(3)
        s0.reset(); s0.prefill([1, 400]); s0.feed(305)
        full = s1.read_logits()
        assert full.shape[0] == world * eng.slots[1].read_logits().shape[0]      # one shard per rank, concatenated
        assert np.array_equal(full[: full.shape[0] // world], eng.slots[1].read_logits())
        # batched steps (TPBatch): launches are mirrored, only the leader collects
        assert lead.batch_capable
        bd = lead.batch
        last = [s0.read_last_token(), s1.read_last_token()]
        h = bd.launch([(0, last[0], s0.n_past), (1, last[1], s1.n_past)])
        h2 = bd.launch_chained()
        t1, t2 = bd.collect(h), bd.collect(h2)
        assert len(t1) == len(t2) == 2
        row = bd.logits_row(1)
        assert row.shape[0] == full.shape[0]
        bd.prefill(0, [301, 302, 303], s0.n_past)
        lead.shutdown()
    else:
        follower_loop(eng, dist)
    np.savez(os.path.join(out_dir, f"serve{rank}.npz"), past=[s.n_past for s in eng.slots], last=[s.read_last_token() for s in eng.slots])
    dist.destroy_process_group()


def test_tp_serving_followers_mirror_the_leader(model_dir, tmp_path):
    """rank 0's scheduler calls reach every rank in the same order (tp_serve.py); gloo, 2 processes, stand-in engines"""
    import torch.multiprocessing as mp
    from ggufb200 import synth
    path = os.path.join(model_dir, "tpserve-tiny.gguf")
    if not os.path.exists(path):
        synth.write_gguf(path, "tiny", "Q4_K_M", seed=0xB200)
    mp.spawn(_tp_serve_worker, args=(2, _free_port(), path, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "serve0.npz"), np.load(tmp_path / "serve1.npz")
    assert list(a["past"]) == list(b["past"]) == [8, 9]
    assert list(a["last"]) == list(b["last"])
