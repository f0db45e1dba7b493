"""GPU (-m gpu): the reference's OWN call sites, executed verbatim against valley_b200.

tests/golden/ref_caller_bodies.json holds the text of four functions of the reference, extracted by
oracle/make_caller_fixture.py (file, line range and SHA-256 recorded):

    valley/serve/model_worker.py:51-94     load_model                         from_pretrained, model.model.multi_image = ...,
                                                                              vision_tower.to(device='cuda', dtype=fp16), model.cuda()
    valley/serve/model_worker.py:320-426   ModelWorker.generate_video_stream  the per-token serving loop (tuple-style cache access)
    valley/inference/run_valley.py:13-18   init_vision_token
    valley/inference/run_valley.py:20-57   main                               model.to(device), model.eval(), model.completion(path)

Each body is exec()'d UNMODIFIED in a namespace whose ``ValleyLlamaForCausalLM`` is the valley_b200 class -- the one import a
maintainer swaps -- and whose unrelated third parties (AutoTokenizer, CLIPImageProcessor, logger, decord) are
small fakes.  What they produce is compared with valley_b200's device-side loops on the same request."""
import hashlib
import json
import os
import re
import types

import numpy as np
import pytest
import torch

import helpers as Hh
from valley_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _bodies():
    d = json.load(open(os.path.join(GOLD, "ref_caller_bodies.json")))
    for k, v in d.items():
        assert hashlib.sha256(v["source"].encode()).hexdigest() == v["sha256"], k       # the fixture is what the script wrote
    return d


class WordTokenizer:
    """Deterministic word-level tokenizer over the model's id space with the HF surface the callers use: ``__call__`` on a
    string or a list (``padding=True``), ``decode`` / ``batch_decode``, ``add_tokens``, ``convert_tokens_to_ids``, ``__len__``.
    Sentinel strings map to the six highest ids; ``stop_id`` decodes to '###' (the separator the reference stops on)."""
    eos_token_id = 2
    pad_token_id = 0
    padding_side = "right"

    def __init__(self, spec, stop_id=None):
        t = syn.sentinel_ids(spec)
        self.V = spec.vocab_size
        self.special = {"<im_patch>": t["im_patch_token"], "<im_start>": t["im_start_token"], "<im_end>": t["im_end_token"],
                        "<vi_frame>": t["vi_frame_token"], "<vi_start>": t["vi_start_token"], "<vi_end>": t["vi_end_token"]}
        self.stop_id = stop_id
        self.rx = re.compile(r"<[a-z_]+>|###|w\d+|[A-Za-z']+|[^\sA-Za-z]")
        self.added = []

    def _word(self, w):
        if w in self.special:
            return self.special[w]
        if re.fullmatch(r"w\d+", w):
            return int(w[1:])
        return 3 + int(hashlib.md5(w.encode()).hexdigest(), 16) % (self.V - 8 - 3 - 1)

    def _ids(self, text):
        return [1] + [self._word(w) for w in self.rx.findall(text)]

    def __call__(self, text, padding=False):
        if isinstance(text, str):
            return types.SimpleNamespace(input_ids=self._ids(text))
        rows = [self._ids(t) for t in text]
        n = max(len(r) for r in rows)
        left = self.padding_side == "left"
        ids = [([self.pad_token_id] * (n - len(r)) + r) if left else (r + [self.pad_token_id] * (n - len(r))) for r in rows]
        am = [([0] * (n - len(r)) + [1] * len(r)) if left else ([1] * len(r) + [0] * (n - len(r))) for r in rows]
        return types.SimpleNamespace(input_ids=ids, attention_mask=am)

    def decode(self, ids, skip_special_tokens=True):
        return "".join(" ###" if int(i) == self.stop_id else f" w{int(i)}" for i in ids if int(i) not in (0, 1, 2))

    def batch_decode(self, rows, skip_special_tokens=True):
        return [self.decode(r.tolist() if torch.is_tensor(r) else r, skip_special_tokens) for r in rows]

    def add_tokens(self, toks, special_tokens=False):
        new = [t for t in toks if t not in self.special and t not in self.added]
        self.added += new
        return len(new)

    def convert_tokens_to_ids(self, toks):
        return self.special[toks] if isinstance(toks, str) else [self.special[t] for t in toks]

    def __len__(self):
        return self.V


class ClipReader:
    """decord.VideoReader surface over an in-memory uint8 clip."""

    def __init__(self, frames):
        self.f = frames

    def __len__(self):
        return len(self.f)

    def get_batch(self, idx):
        return torch.from_numpy(self.f[np.asarray(idx)])

    def get_avg_fps(self):
        return 25.0


def _checkpoint(tmp_path, spec, sd):
    from test_host_logic import _write_checkpoint
    d = tmp_path / "valley-tiny-ckpt"
    d.mkdir()
    _write_checkpoint(str(d), spec, {k: v.bfloat16() for k, v in sd.items()}, "safetensors")
    return str(d)


def _namespace(tok, **extra):
    """What the reference scripts import at module level, with the ONE swapped import."""
    from valley_b200 import model as M
    log = types.SimpleNamespace(info=lambda *a, **k: None, warning=lambda *a, **k: None)
    ns = dict(torch=torch, os=os, json=json, np=np, logger=log,
              ValleyLlamaForCausalLM=M.ValleyLlamaForCausalLM,                                  # <- the swapped import
              AutoTokenizer=types.SimpleNamespace(from_pretrained=lambda *a, **k: tok),
              CLIPImageProcessor=types.SimpleNamespace(from_pretrained=lambda *a, **k: object()),
              disable_torch_init=lambda: None,
              DEFAULT_IMAGE_PATCH_TOKEN=M.DEFAULT_IMAGE_PATCH_TOKEN, DEFAULT_IM_START_TOKEN=M.DEFAULT_IM_START_TOKEN,
              DEFAULT_IM_END_TOKEN=M.DEFAULT_IM_END_TOKEN, DEFAULT_VIDEO_FRAME_TOKEN=M.DEFAULT_VIDEO_FRAME_TOKEN,
              DEFAULT_VI_START_TOKEN=M.DEFAULT_VI_START_TOKEN, DEFAULT_VI_END_TOKEN=M.DEFAULT_VI_END_TOKEN,
              DEFAULT_VIDEO_TOKEN="<video>", DEFAULT_IMAGE_TOKEN="<image>")
    ns.update(extra)
    return ns


def test_model_worker_load_model_and_serving_loop_run_unmodified(tmp_path):
    """model_worker.py:51-94 + :320-426 on our class: the load path (.to on the vision tower, .cuda() on the model, token-id
    setup) and the per-token loop with its tuple-style cache access; its streamed texts equal valley_b200.serving's."""
    from valley_b200 import serving
    from valley_b200.model import ValleyLlamaForCausalLM
    bodies = _bodies()
    spec = syn.TINY
    sd = Hh.bf16_weights(spec, 0)
    path = _checkpoint(tmp_path, spec, sd)
    tok = WordTokenizer(spec)
    ns = _namespace(tok, args=types.SimpleNamespace(stream_interval=2))
    exec(bodies["valley/serve/model_worker.py:load_model"]["source"], ns)
    tokenizer, model, image_processor, context_len = ns["load_model"](path, "valley-tiny", 1)
    assert isinstance(model, ValleyLlamaForCausalLM) and tokenizer is tok and context_len == 2048
    vc = model.get_model().vision_tower.config
    t = syn.sentinel_ids(spec)
    assert (vc.im_patch_token, vc.im_start_token, vc.im_end_token) == (t["im_patch_token"], t["im_start_token"], t["im_end_token"])
    assert vc.use_im_start_end is True and model.model.multi_image is True and model.device.type == "cuda"

    exec(bodies["valley/serve/model_worker.py:ModelWorker.generate_video_stream"]["source"], ns)
    clip = syn.make_pixels(1, 3, 8)[0].permute(1, 0, 2, 3).contiguous()           # what ModelWorker.load_video returns: [3,T,224,224]
    worker = types.SimpleNamespace(tokenizer=tokenizer, model=model, image_processor=image_processor, is_multimodal=True,
                                   context_len=context_len, load_video=lambda p: clip)
    words = " ".join(f"w{i}" for i in torch.randint(3, 900, (12,), generator=torch.Generator().manual_seed(1)).tolist())
    params = dict(prompt=f"{words} <video> w77 w78", videos=["clip.mp4"], temperature=0.0, max_new_tokens=9, stop=" w1 w1 w1")
    chunks = list(ns["generate_video_stream"](worker, params))
    ref_texts = [json.loads(c[:-1].decode())["text"] for c in chunks]
    assert 1 <= len(ref_texts) <= 5 and all(json.loads(c[:-1].decode())["error_code"] == 0 for c in chunks)     # i = 0, 2, 4, 6, 8
    # the same request through the device loop (serving.generate_stream): identical streamed texts.  The worker never sets the
    # vi_* ids (load_model only sets im_*), so both sides take the reference's silent image-only fallback for the frame tokens.
    mine = [d["text"] for d in serving.generate_stream(model, tokenizer, dict(params, video=clip.permute(1, 0, 2, 3)), stream_interval=2)]
    assert mine == ref_texts


def test_run_valley_main_runs_unmodified_and_completion_takes_a_path(tmp_path):
    """run_valley.py:13-57 on our class: from_pretrained, init_vision_token, model.to(device), model.eval(), then
    model.completion(tokenizer, args.video_file, message, gen_kwargs, device) with a FILE PATH; the reply ends at '###' via the
    KeywordsStoppingCriteria the reference passes, and equals the text derived from a free-running greedy decode."""
    from oracle import preprocess_oracle as P
    from valley_b200 import model as M
    bodies = _bodies()
    spec = syn.TINY
    path = _checkpoint(tmp_path, spec, Hh.bf16_weights(spec, 0))
    frames = np.random.default_rng(5).integers(0, 256, (21, 300, 400, 3), dtype=np.uint8)
    query, system = "Describe this video concisely.\n<video>", "You are Valley."
    message = [{"role": "system", "content": system}, {"role": "user", "content": query}]

    # what free-running greedy decoding emits for main()'s request (the oracle preprocessing feeds the same pixels the device
    # pipeline produces -- bit-exact, test_frame_preprocessing_is_bit_exact); pick the token that will play '###'
    tok0 = WordTokenizer(spec)
    probe = M.ValleyLlamaForCausalLM.from_pretrained(path, torch_dtype=torch.float16)
    for k_, v_ in syn.sentinel_ids(spec).items():
        setattr(probe.get_model().vision_tower.config, k_, v_)
    ids = torch.as_tensor(probe.build_inputs(tok0, message).input_ids)
    px = torch.from_numpy(P.preprocess_frames(frames[P.fixed_frame_indices(21, 8)])).half()[None]
    free = probe.generate(input_ids=ids.cuda(), images=px.cuda(), max_new_tokens=48, eos_token_id=None)[0, ids.shape[1]:].tolist()
    k = next(i for i in range(5, 48) if free[i] not in free[:i] and free[i] not in (0, 1, 2))
    tok = WordTokenizer(spec, stop_id=free[k])            # encoding is unchanged (same prompt ids); only decoding shows '###'
    n_used = next((i for i in range(1, len(free) + 1)      # HF loop: eos (config default 2) or, from the 2nd call on, the keyword
                   if free[i - 1] == 2 or (i >= 2 and "###" in tok.decode(free[:i]))), len(free))
    want = probe.process_response([tok.decode(free[:n_used])])
    del probe

    opened, printed = [], []

    def fake_open(p):
        opened.append(p)
        return ClipReader(frames)

    ns = _namespace(tok, print=lambda *a, **kw: printed.append(a[0] if a else None), PeftModel=None, PeftConfig=None,
                    DEFAULT_SYSTEM=system)
    exec(bodies["valley/inference/run_valley.py:init_vision_token"]["source"], ns)
    exec(bodies["valley/inference/run_valley.py:main"]["source"], ns)
    old = M.ValleyLlamaForCausalLM.video_reader_factory
    M.ValleyLlamaForCausalLM.video_reader_factory = staticmethod(fake_open)
    try:
        ns["main"](types.SimpleNamespace(model_name=path, query=query, video_file="some/clip.mp4", vision_tower=None, system_prompt=""))
    finally:
        M.ValleyLlamaForCausalLM.video_reader_factory = old
    assert opened == ["some/clip.mp4"]
    response = printed[-1]
    assert response == want, (response, want)
    assert isinstance(response, list) and len(response) == 1 and "###" not in response[0] and n_used <= k + 1


def test_module_surface_validates_instead_of_silently_accepting():
    """.to()/.cuda()/.half()/.eval() return self for what the reference's callers pass; a CPU move or fp32 cast raises."""
    from valley_b200 import _lib
    spec = syn.TINY
    m = Hh.build_model(spec, Hh.bf16_weights(spec, 0))
    vt = m.get_model().vision_tower
    assert m.to(torch.device("cuda")) is m and m.cuda() is m and m.half() is m and m.eval() is m and m.to("cuda:0") is m
    assert vt.to(device="cuda", dtype=torch.float16) is vt and vt.to(torch.device("cuda"), dtype=torch.float16) is vt
    assert vt.device == m.device and m.dtype == torch.bfloat16 and m.training is False
    for bad in (lambda: m.to("cpu"), lambda: m.to(torch.float32), lambda: vt.to(device="cpu"), lambda: m.train()):
        with pytest.raises(_lib.VlyError):
            bad()
    if torch.cuda.device_count() > 1:
        with pytest.raises(_lib.VlyError):
            m.to("cuda:1")
    with pytest.raises(NotImplementedError):
        m(input_ids=syn.make_prompt_ids(spec, 1, 2, 0).cuda(), output_hidden_states=True)
    with pytest.raises(_lib.VlyError):
        m.get_model().mm_projector.weight


def test_generate_hf_defaults_eos_from_config_and_pad_mask_inference():
    """HF generate defaults the reference relies on: eos comes from config.eos_token_id, finished rows are padded
    in the host-visible loop too, and a prompt containing pad_token_id gets its attention_mask inferred."""
    from valley_b200.model import KeywordsStoppingCriteria
    spec = syn.TINY
    m = Hh.build_model(spec, Hh.bf16_weights(spec, 0))
    B, n = 3, 10
    ids, px = syn.make_prompt_ids(spec, B, 2, 6), syn.make_pixels(B, 2, 6)
    S = ids.shape[1]
    free = m.generate(input_ids=ids.cuda(), images=px.cuda(), max_new_tokens=n)[:, S:].cpu()
    eos = int(free[0, 3])
    m.config.eos_token_id, m.config.pad_token_id = eos, None
    try:
        got = m.generate(input_ids=ids.cuda(), images=px.cuda(), max_new_tokens=n)[:, S:].cpu()          # device loop, eos from config
        never = lambda seq, scores: False
        host = m.generate(input_ids=ids.cuda(), images=px.cuda(), max_new_tokens=n, stopping_criteria=[never])[:, S:].cpu()
        assert torch.equal(got, host)                                                                   # host loop pads finished rows too
        for b in range(B):
            hit = (free[b] == eos).nonzero()
            if len(hit) and int(hit[0]) + 1 < got.shape[1]:
                assert (got[b, int(hit[0]) + 1:] == eos).all()                                          # HF: pad defaults to eos
        full = m.generate(input_ids=ids.cuda(), images=px.cuda(), max_new_tokens=n, eos_token_id=None)[:, S:].cpu()
        assert torch.equal(full, free)                                                                  # explicit None = run to length
    finally:
        m.config.eos_token_id = None
    # pad-mask inference: left-padded ids + pad_token_id and no attention_mask == the same call with the explicit mask
    P = 5
    pad = torch.zeros(B, P, dtype=torch.int64)
    ids_p = torch.cat([pad, ids], 1)
    am = torch.cat([torch.zeros(B, P, dtype=torch.int64), torch.ones_like(ids)], 1)
    a = m.generate(input_ids=ids_p.cuda(), images=px.cuda(), max_new_tokens=4, pad_token_id=0)
    b = m.generate(input_ids=ids_p.cuda(), images=px.cuda(), max_new_tokens=4, attention_mask=am.cuda())
    assert torch.equal(a, b)
    # KeywordsStoppingCriteria (data_util.py:40-56): first call records the prompt length, later calls test the decoded tail
    tk = WordTokenizer(spec, stop_id=int(free[0, 4]))
    crit = KeywordsStoppingCriteria(["###"], tk, ids[:1])
    out = m.generate(input_ids=ids[:1].cuda(), images=px[:1].cuda(), max_new_tokens=n, stopping_criteria=[crit])
    assert out.shape[1] == S + 5 and int(out[0, -1]) == int(free[0, 4])
