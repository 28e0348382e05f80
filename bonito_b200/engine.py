"""
Plan builder + executor for the LSTM-CRF encoder (the `bonito.crf` fast / hac models; LSTM widths 96, 128, 256, 384).

`compile_lstm_crf(encoder)` walks a `bonito_b200.nn` module tree of the shape the reference's
configs describe (`/root/reference/bonito/models/configs/dna_r10.4.1@v4.3.toml`,
`bonito/crf/model.py:150-162`):

    Convolution(1->C1,k5) , Convolution(C1->C2,k5) , Convolution(C2->H,kW,stride s)
    Permute([2,0,1]) , LSTM x L (alternating reverse) , LinearCRFEncoder [, Clamp]

(LinearCRFEncoder: a plain linear head followed by a Clamp layer, as in the v4+ configs, or the old-style head with
activation = "tanh" and / or a scale and no Clamp; fixed blank_score) and packs the weights into the operand layouts of
the sm_100a kernels (include/bonito_b200.h).  This is the native swap-in the reference performs in `Model.use_koi`
(`bonito/crf/model.py:240-246`, koi.lstm.update_graph); like koi it returns scores as `[N, T, C]` fp16 without the blank
column.  Anything else raises `UnsupportedModel`.

Width 384 (hac), the headline path (`forward_tiles`, DESIGN.md sections 3 and 4): activations tile-major
`[tile][T][48][H]`, gate pre-activations `[tile][T][6][48][256]`; per layer ONE input GEMM over all tiles and ONE launch
of the recurrent kernel (one 6-CTA cluster per 48-chunk tile, 11 clusters for 512 chunks), 15 launches per batch, issued by
one C call (`b200_lstm_crf_fwd`) unless per-kernel events or intermediate activations are asked for.  Buffers are cached
per (batch, chunk length, slot): `slot` selects one of several independent buffer sets, so that consecutive batches can
be in flight on different streams (`score_batches`, bench.py).

Other widths (`forward_tiled`): the generic `[T][N][4H]` layout and the `mma.sync` recurrent kernel, tiles of 32 chunks
pipelined on per-tile streams; `B200_LSTM_TILE=0` sends width 384 down this path with the first-generation tcgen05 kernel
(8-CTA clusters), `B200_TILE_STREAMS=1` gives the tile-layout path per-tile streams as well (the round-1 schedule).
"""

import torch

from bonito_b200 import native
from bonito_b200 import nn as bnn

_ACT_CODES = {None: native.ACT_NONE, "swish": native.ACT_SWISH, "tanh": native.ACT_TANH}


class UnsupportedModel(NotImplementedError):
    pass


def _act_code(module):
    if module is None:
        return native.ACT_NONE
    name = getattr(module, "name", None)
    if name not in _ACT_CODES:
        raise UnsupportedModel(f"activation {module!r} has no native kernel")
    return _ACT_CODES[name]


def _folded_conv(layer):
    """(weight, bias) of a Convolution with any BatchNorm folded in (bonito/nn.py:447-454)."""
    conv = layer.conv
    if layer.norm is not None:
        if not isinstance(layer.norm, bnn.BatchNorm):
            raise UnsupportedModel(f"norm {layer.norm!r} has no native kernel")
        conv = torch.nn.utils.fusion.fuse_conv_bn_eval(conv.eval(), layer.norm.bn.eval())
    return conv.weight.detach(), None if conv.bias is None else conv.bias.detach()


def _dev16(t, device):
    return None if t is None else t.to(device=device, dtype=torch.float16).contiguous()


class _Stage:
    """Context manager recording a CUDA event pair around one kernel launch (no-op without a sink)."""

    def __init__(self, name, sink):
        self.name, self.sink = name, sink

    def __enter__(self):
        if self.sink is not None:
            self.start = torch.cuda.Event(enable_timing=True)
            self.start.record()

    def __exit__(self, *exc):
        if self.sink is not None:
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            self.sink.append((self.name, self.start, end))
        return False


class _StreamStage(_Stage):
    """`_Stage` whose events are recorded on an explicit stream."""

    def __init__(self, name, sink, stream):
        super().__init__(name, sink)
        self.stream = stream

    def __enter__(self):
        if self.sink is not None:
            self.start = torch.cuda.Event(enable_timing=True)
            self.start.record(self.stream)

    def __exit__(self, *exc):
        if self.sink is not None:
            end = torch.cuda.Event(enable_timing=True)
            end.record(self.stream)
            self.sink.append((self.name, self.start, end))
        return False


class LstmCrfPlan:
    """Packed weights + cached buffers for one LSTM-CRF encoder on one device."""

    def __init__(self, encoder, device, quantize=False):
        layers = list(encoder.children())
        convs = [m for m in layers if isinstance(m, bnn.Convolution)]
        lstms = [m for m in layers if isinstance(m, bnn.LSTM)]
        crfs = [m for m in layers if isinstance(m, bnn.LinearCRFEncoder)]
        clamps = [m for m in layers if isinstance(m, bnn.Clamp)]
        others = [m for m in layers if not isinstance(
            m, (bnn.Convolution, bnn.LSTM, bnn.LinearCRFEncoder, bnn.Clamp, bnn.Permute))]
        if len(convs) != 3 or not lstms or len(crfs) != 1 or others or len(clamps) > 1:
            raise UnsupportedModel("native LSTM-CRF path needs 3 convolutions, >=1 LSTM, one LinearCRFEncoder "
                                   f"and at most one Clamp; got {[type(m).__name__ for m in layers]}")
        self.device = torch.device(device)

        # --- conv stem (conv1 + conv2) -------------------------------------------------------
        c1, c2, c3 = convs
        for c in (c1, c2):
            k = c.conv.kernel_size[0]
            if c.conv.stride[0] != 1 or c.conv.padding[0] != k // 2 or k % 2 == 0:
                raise UnsupportedModel("conv stem layers must be stride 1 with 'same' padding")
        if c1.conv.in_channels != 1:
            raise UnsupportedModel("the encoder must take a single input feature")
        w1, b1 = _folded_conv(c1)
        w2, b2 = _folded_conv(c2)
        self.w1, self.b1, self.act1 = _dev16(w1, device), _dev16(b1, device), _act_code(c1.activation)
        self.w2, self.b2, self.act2 = _dev16(w2, device), _dev16(b2, device), _act_code(c2.activation)

        # --- strided conv as GEMM: weight [H][C2][K3] -> [H][K3*C2], k = tap*C2 + cin -------
        w3, b3 = _folded_conv(c3)
        self.hidden, self.c2, self.k3 = w3.shape
        self.s3, self.pad3 = c3.conv.stride[0], c3.conv.padding[0]
        self.w3 = _dev16(w3.permute(0, 2, 1).reshape(self.hidden, -1), device)
        self.b3, self.act3 = _dev16(b3, device), _act_code(c3.activation)
        if (self.k3 * self.c2) % 8 or (self.s3 * self.c2) % 8:
            raise UnsupportedModel("strided conv window/stride must be multiples of 8 elements")

        # --- LSTM stack ------------------------------------------------------------------------
        H = self.hidden
        if native.lstm_cluster_size(H) == 0:
            raise UnsupportedModel(f"LSTM hidden size {H} has no native kernel")
        # H = 384: second-generation recurrent kernel (48-chunk tiles, 6-CTA clusters, gx streamed through shared memory)
        self.tile = native.lstm_tile_chunks(H)            # 0: only the generic-layout kernel exists for this width
        self.tile_cs = native.lstm_tile_cluster(H)
        unit = torch.arange(H)
        perm_ih = (torch.arange(4)[None, :] * H + unit[:, None]).reshape(-1)          # [unit][gate]
        perm_hh = (torch.arange(H // 8)[:, None, None] * 8 + torch.arange(4)[None, :, None] * H
                   + torch.arange(8)[None, None, :]).reshape(-1)                      # [unit/8][gate][unit%8]
        self.lstm = []
        for m in lstms:
            r = m.rnn
            if r.hidden_size != H or r.input_size != H or r.num_layers != 1 or r.bidirectional:
                raise UnsupportedModel("native LSTM path needs single-layer unidirectional LSTMs of equal width")
            bias = torch.zeros(4 * H, dtype=torch.float32)
            if r.bias:
                bias = r.bias_ih_l0.detach().float().cpu() + r.bias_hh_l0.detach().float().cpu()
            self.lstm.append(dict(
                wih=_dev16(r.weight_ih_l0.detach().cpu()[perm_ih], device),
                bias=_dev16(bias[perm_ih], device),
                whh=_dev16(r.weight_hh_l0.detach().cpu()[perm_hh], device),
                reverse=bool(m.reverse),
            ))
            if quantize:
                # --quantize (reference: koi's int8 LSTM, bonito/crf/model.py:245): the input projection runs on int8 tensor
                # cores.  Weights: symmetric per-row int8; activations: LSTM inputs live in (-1, 1) (tanh / o*tanh(c)), fixed
                # scale 127.  gx = (q_x . q_w) * s_w / 127 + bias.  The recurrent weights and the h exchange stay fp16.
                w = r.weight_ih_l0.detach().float().cpu()[perm_ih]
                s_w = w.abs().amax(dim=1).clamp(min=1e-8) / 127.0
                self.lstm[-1]["wih_q"] = torch.round(w / s_w[:, None]).clamp(-127, 127).to(torch.int8).to(device).contiguous()
                self.lstm[-1]["wih_scale"] = (s_w / 127.0).to(device=device, dtype=torch.float32).contiguous()
        self.quantize = bool(quantize)
        if self.quantize and not (self.tile and H % 16 == 0):
            raise UnsupportedModel("the int8 input projection (--quantize) needs the tile-layout LSTM path (hidden size 384)")

        # --- linear CRF head (+ clamp) ---------------------------------------------------------
        crf = crfs[0]
        if crf.permute is not None:
            raise UnsupportedModel("native LinearCRFEncoder supports permute=None")
        crf_act = _act_code(crf.activation)
        if crf_act not in (native.ACT_NONE, native.ACT_TANH) or ((crf_act != native.ACT_NONE or crf.scale is not None) and clamps):
            raise UnsupportedModel("native LinearCRFEncoder supports tanh and / or a scale (old-style configs), or a Clamp layer "
                                   "behind a plain linear head (v4+ configs)")
        if crf.blank_score is None:
            raise UnsupportedModel("native decode needs a fixed blank_score")
        self.n_base, self.state_len, self.blank_score = crf.n_base, crf.state_len, float(crf.blank_score)
        self.wl = _dev16(crf.linear.weight.detach(), device)
        self.bl = _dev16(None if crf.linear.bias is None else crf.linear.bias.detach(), device)
        self.n_scores = self.wl.shape[0]
        if clamps:
            self.act_l, self.lo, self.hi = native.ACT_CLAMP, float(clamps[0].min), float(clamps[0].max)
        elif crf_act == native.ACT_TANH and crf.scale is not None:     # e.g. the dna_r9.4.1 configs: tanh, scale 5.0
            self.act_l, self.lo, self.hi = native.ACT_TANH_SCALE, float(crf.scale), 0.0
        elif crf.scale is not None:
            self.act_l, self.lo, self.hi = native.ACT_SCALE, float(crf.scale), 0.0
        else:
            self.act_l, self.lo, self.hi = crf_act, 0.0, 0.0
        self._bufs = {}

    # ------------------------------------------------------------------------------------------
    def frames(self, L):
        return (L + 2 * self.pad3 - self.k3) // self.s3 + 1

    def _buffers(self, N, L):
        key = (N, L)
        if key not in self._bufs:
            self._bufs.clear()
            T = self.frames(L)
            need = max(self.pad3 + L, (T - 1) * self.s3 + self.k3)
            Tp = -(-need // self.s3)
            Lp = Tp * self.s3
            dev, f16 = self.device, torch.float16
            tail = self.k3 * self.c2  # the last window of the overlapping-row view reads past row N*Tp-1
            self._bufs[key] = dict(
                T=T, Tp=Tp, Lp=Lp,
                stem=torch.empty(N * Lp * self.c2 + tail, dtype=f16, device=dev),
                ya=torch.empty(T, N, self.hidden, dtype=f16, device=dev),
                yb=torch.empty(T, N, self.hidden, dtype=f16, device=dev),
                gx=torch.empty(T, N, 4 * self.hidden, dtype=f16, device=dev),
            )
            self._bufs[key]["stem"][-tail:].zero_()
        return self._bufs[key]

    @property
    def supports_slots(self):
        """Independent buffer sets (several batches in flight) exist for the tile-layout path only."""
        import os
        return bool(self.tile) and os.environ.get("B200_LSTM_TILE", "1") != "0"

    TILE = 32  # chunks per recurrent cluster
    # CTAs a per-tile GEMM may occupy while recurrent clusters of other tiles are resident (0 = all SMs)
    TILE_GEMM_CTAS = 0

    def _tile_buffers(self, N, L):
        key = ("tiled", N, L)
        if key not in self._bufs:
            self._bufs.clear()
            T = self.frames(L)
            need = max(self.pad3 + L, (T - 1) * self.s3 + self.k3)
            Tp = -(-need // self.s3)
            Lp = Tp * self.s3
            dev, f16, H = self.device, torch.float16, self.hidden
            nt = -(-N // self.TILE)
            tail = self.k3 * self.c2
            stem = torch.empty(N * Lp * self.c2 + tail, dtype=f16, device=dev)
            stem[-tail:].zero_()
            self._bufs[key] = dict(
                T=T, Tp=Tp, Lp=Lp, nt=nt, stem=stem,
                ya=torch.empty(nt, T, self.TILE, H, dtype=f16, device=dev),
                yb=torch.empty(nt, T, self.TILE, H, dtype=f16, device=dev),
                gx=torch.empty(nt, T, self.TILE, 4 * H, dtype=f16, device=dev),
                streams=_LazyStreams(dev, nt), rec_streams=_LazyStreams(dev, nt),
                rec_ready=[torch.cuda.Event() for _ in range(nt)], rec_done=[torch.cuda.Event() for _ in range(nt)],
                done=[torch.cuda.Event() for _ in range(nt)],
                head=[torch.cuda.Event() for _ in range(nt)],
                start=torch.cuda.Event(),
            )
        return self._bufs[key]

    def forward_tiled(self, x, out=None, gemm_impl=native.GEMM_AUTO, events=None, decode=None):
        """
        Tile-pipelined forward (see module docstring): same result as `forward(..., tiled=False)`.
        `decode=(qscale, qbias)`: also enqueue the CRF decode of every tile on that tile's stream, right behind its
        CRF GEMM (it then overlaps the recurrences of the other tiles); the results are parked in `DECODE_CACHE` and
        handed out by `CrfDecoder` when it is asked to decode exactly these scores with exactly these parameters.
        Off by default: measured 5 % slower end to end (34.9 vs 33.0 ms/step), the decode CTAs land on the SMs of
        the recurrent clusters and lengthen their per-step critical path.
        """
        if x.dim() == 3:
            x = x[:, 0, :]
        x = x.to(device=self.device, dtype=torch.float16).contiguous()
        N, L = x.shape
        H, TB = self.hidden, self.TILE
        b = self._tile_buffers(N, L)
        T, Tp, Lp, nt = b["T"], b["Tp"], b["Lp"], b["nt"]
        if out is None:
            out = torch.empty(N, T, self.n_scores, dtype=torch.float16, device=self.device)
        main = torch.cuda.current_stream()
        dec = None
        if decode is not None:
            ws_tile = native.crf_decode_workspace_bytes(TB, T, self.state_len)
            if b.get("dec_ws") is None or b["dec_ws"].numel() < ws_tile * nt:
                b["dec_ws"] = torch.empty(ws_tile * nt, dtype=torch.uint8, device=self.device)
            dec = [torch.empty(N, T, dtype=torch.uint8, device=self.device) for _ in range(3)]  # moves, seq, qual

        with _Stage("conv_stem", events):
            native.conv_stem(x, self.w1, self.b1, self.act1, self.w2, self.b2, self.act2, b["stem"], Lp, self.pad3)
        tiles = []
        for i in range(nt):
            n0 = i * TB
            nb = min(TB, N - n0)
            tiles.append((i, n0, nb, b["streams"][i]))

        def staged(name, st):
            return _StreamStage(name, events, st)

        import os
        cap = int(os.environ.get("B200_TILE_GEMM_CTAS", self.TILE_GEMM_CTAS))
        stagger = os.environ.get("B200_TILE_STAGGER", "0") != "0"
        rec_prio = os.environ.get("B200_LSTM_PRIO", "1") != "0"   # recurrent kernels on high-priority side streams

        # Head of the pipeline.  With every tile's first GEMMs on its own stream they share the machine and finish
        # together; the 15 cluster slots then fill and drain in lock-step and the GEMMs of the next layer again arrive
        # all at once (measured: 4.2 ms per layer = 1.2 ms GEMM phase + 2.5 ms recurrence + the 16th tile trailing).
        # B200_TILE_STAGGER=1 serialises the head GEMMs on the main stream so that the tiles stay staggered; measured
        # slightly slower (29.4 vs 28.8 ms/step): persistent GEMM CTAs then squat on SMs a waiting cluster needs, and
        # every recurrent launch queues ~0.9 ms for 8 free SMs inside one GPC.  Lock-step is the default.
        # The recurrent kernels go to a HIGH-PRIORITY side stream per tile (B200_LSTM_PRIO=0 turns that off): when SMs free
        # up, a waiting cluster is placed before the queued CTAs of the other tiles' GEMMs (24.2 -> 22.9 ms/step).
        first = self.lstm[0]
        for i, n0, nb, st in tiles:
            head = main if stagger else st
            if not stagger and i == 0:
                b["start"].record(main)
            if not stagger:
                st.wait_event(b["start"])
            with staged("conv_gemm", head):   # rows r = i_chunk*Tp + t of this tile -> ya[tile][t][i_chunk]
                native.gemm(b["stem"][n0 * Lp * self.c2:], self.s3 * self.c2, self.w3, self.b3, b["ya"][i], H, nb * Tp, H,
                            self.k3 * self.c2, act=self.act3, rows_inner=Tp, valid_inner=T, stride_inner=nb,
                            stride_outer=1, impl=gemm_impl, stream=head)
            with staged("lstm_in_gemm", head):
                native.gemm(b["ya"][i], H, first["wih"], first["bias"], b["gx"][i], 4 * H, T * nb, 4 * H, H,
                            impl=gemm_impl, stream=head)
            if stagger:
                b["head"][i].record(main)
                st.wait_event(b["head"][i])
        cur, nxt = b["ya"], b["yb"]
        for li, layer in enumerate(self.lstm):
            for i, n0, nb, st in tiles:
                if li > 0:
                    with staged("lstm_in_gemm", st):
                        native.gemm(cur[i], H, layer["wih"], layer["bias"], b["gx"][i], 4 * H, T * nb, 4 * H, H,
                                    impl=gemm_impl, stream=st, max_ctas=cap)
                if rec_prio:
                    rs = b["rec_streams"][i]
                    b["rec_ready"][i].record(st)
                    rs.wait_event(b["rec_ready"][i])
                    with staged("lstm_rec", rs):
                        native.lstm_rec(b["gx"][i], layer["whh"], nxt[i], T, nb, H, layer["reverse"], stream=rs)
                    b["rec_done"][i].record(rs)
                    st.wait_event(b["rec_done"][i])
                else:
                    with staged("lstm_rec", st):
                        native.lstm_rec(b["gx"][i], layer["whh"], nxt[i], T, nb, H, layer["reverse"], stream=st)
            cur, nxt = nxt, cur
        for i, n0, nb, st in tiles:
            with staged("crf_gemm", st):    # rows r = t*nb + i_chunk -> out[n0 + i_chunk][t]
                native.gemm(cur[i], H, self.wl, self.bl, out[n0:], self.n_scores, T * nb, self.n_scores, H,
                            act=self.act_l, lo=self.lo, hi=self.hi, rows_inner=nb, valid_inner=nb, stride_inner=T,
                            stride_outer=1, impl=gemm_impl, stream=st, max_ctas=cap)
            if dec is not None:
                with staged("crf_decode", st):
                    native.crf_decode(out[n0:n0 + nb], self.state_len, self.blank_score, decode[0], decode[1],
                                      b["dec_ws"][i * ws_tile:], dec[0][n0:], dec[1][n0:], dec[2][n0:], stream=st)
            b["done"][i].record(st)
        for i in range(nt):
            main.wait_event(b["done"][i])
        if dec is not None:
            DECODE_CACHE.put(out, (self.state_len, self.blank_score, float(decode[0]), float(decode[1])), tuple(dec))
        return out

    # ------------------------------------------------------------------------------------------
    # tile layout (H = 384): activations [tile][T][48][H], gate pre-activations [tile][T][6][48][256]
    # ------------------------------------------------------------------------------------------
    def _tile_layout_buffers(self, N, L, slot=0):
        key = ("tiles", N, L, slot)
        if key not in self._bufs:
            for k in [k for k in self._bufs if k[0] != "tiles" or k[1:3] != (N, L)]:
                del self._bufs[k]
            T = self.frames(L)
            need = max(self.pad3 + L, (T - 1) * self.s3 + self.k3)
            Tp = -(-need // self.s3)
            Lp = Tp * self.s3
            dev, f16, H, TB = self.device, torch.float16, self.hidden, self.tile
            nt = -(-N // TB)
            tail = self.k3 * self.c2
            stem = torch.empty(N * Lp * self.c2 + tail, dtype=f16, device=dev)
            stem[-tail:].zero_()
            self._bufs[key] = dict(
                T=T, Tp=Tp, Lp=Lp, nt=nt, stem=stem,
                # zero-filled once: rows of chunks beyond the batch (last tile) are never written and must stay finite
                ya=torch.zeros(nt, T, TB, H, dtype=f16, device=dev),
                yb=torch.zeros(nt, T, TB, H, dtype=f16, device=dev),
                gx=torch.zeros(nt, T, self.tile_cs, TB, 4 * H // self.tile_cs, dtype=f16, device=dev),
                yq=torch.empty(nt * T * TB * H, dtype=torch.int8, device=dev) if self.quantize else None,
                # staging of the recurrent kernel's h all-gather: one region per tile (tiles run concurrently)
                hx=torch.empty(nt, native.lstm_rec_tile_workspace_bytes(TB), dtype=torch.uint8, device=dev),
                streams=_LazyStreams(dev, nt), rec_streams=_LazyStreams(dev, nt),
                rec_ready=[torch.cuda.Event() for _ in range(nt)], rec_done=[torch.cuda.Event() for _ in range(nt)],
                done=[torch.cuda.Event() for _ in range(nt)],
                start=torch.cuda.Event(),
            )
        return self._bufs[key]

    def _plan_struct(self, b, N, L):
        """`b200_lstm_crf_plan` for this geometry and buffer set (cached in the buffer dict)."""
        if "struct" not in b:
            p = native.LstmCrfPlanStruct()
            p.n, p.l, p.t, p.tp = N, L, b["T"], b["Tp"]
            p.c1, _, p.k1 = self.w1.shape
            p.c2, _, p.k2 = self.w2.shape
            p.act1, p.act2 = self.act1, self.act2
            p.hidden, p.k3, p.s3, p.pad3, p.act3 = self.hidden, self.k3, self.s3, self.pad3, self.act3
            p.n_lstm, p.n_scores, p.act_l, p.lo, p.hi = len(self.lstm), self.n_scores, self.act_l, self.lo, self.hi
            ptr = lambda t: None if t is None else t.data_ptr()
            p.w1, p.b1, p.w2, p.b2, p.w3, p.b3 = ptr(self.w1), ptr(self.b1), ptr(self.w2), ptr(self.b2), ptr(self.w3), ptr(self.b3)
            p.wl, p.bl = ptr(self.wl), ptr(self.bl)
            for i, layer in enumerate(self.lstm):
                p.reverse[i] = int(layer["reverse"])
                p.wih[i], p.bias[i], p.whh[i] = ptr(layer["wih"]), ptr(layer["bias"]), ptr(layer["whh"])
            p.stem, p.ya, p.yb, p.gx, p.hx = (ptr(b[k]) for k in ("stem", "ya", "yb", "gx", "hx"))
            b["struct"] = p
        return b["struct"]

    def forward_tiles(self, x, out=None, gemm_impl=native.GEMM_AUTO, events=None, return_features=False, streams=False,
                      slot=0):
        """
        Forward in the tile layout.  `streams=False`: layer by layer on the current stream -- one input-projection GEMM
        and ONE recurrent launch (a cluster per tile, all in one wave) per layer.  `streams=True`: every tile gets its own
        stream (conv GEMM -> 5 x (input GEMM -> recurrent cluster) -> CRF GEMM), recurrent launches on high-priority side
        streams, so the GEMMs of one tile fill the SMs the other tiles' clusters leave free.  Bit-identical results.
        `slot` selects one of several independent buffer sets (batches in flight at the same time).
        """
        import os
        if x.dim() == 3:
            x = x[:, 0, :]
        x = x.to(device=self.device, dtype=torch.float16).contiguous()
        N, L = x.shape
        H, TB, CS = self.hidden, self.tile, self.tile_cs
        CW = 4 * H // CS                     # gx columns per cluster rank (256)
        b = self._tile_layout_buffers(N, L, slot)
        T, Tp, Lp, nt = b["T"], b["Tp"], b["Lp"], b["nt"]
        if out is None:
            out = torch.empty(N, T, self.n_scores, dtype=torch.float16, device=self.device)
        main = torch.cuda.current_stream()
        feats = {}
        cap = int(os.environ.get("B200_TILE_GEMM_CTAS", self.TILE_GEMM_CTAS))

        def staged(name, st):
            return _StreamStage(name, events, st)

        def conv_gemm(i, st):       # rows r = i_chunk*Tp + t of tile i -> ya[i][t][i_chunk]
            n0 = i * TB
            nb = min(TB, N - n0)
            with staged("conv_gemm", st):
                native.gemm(b["stem"][n0 * Lp * self.c2:], self.s3 * self.c2, self.w3, self.b3, b["ya"][i], H, nb * Tp, H,
                            self.k3 * self.c2, act=self.act3, rows_inner=Tp, valid_inner=T, stride_inner=TB,
                            stride_outer=1, impl=gemm_impl, stream=st)

        def in_gemm(src, layer, tiles, st, max_ctas=0):      # rows (tile, t, chunk) -> gx[tile][t][rank][chunk][CW]
            i0, cnt = tiles
            if self.quantize:
                rows = cnt * T * TB
                yq = b["yq"][i0 * T * TB * H:]
                with staged("quantize_i8", st):
                    native.quantize_i8(src[i0:i0 + cnt], yq[:rows * H], 127.0, stream=st)
                with staged("lstm_in_gemm", st):
                    native.gemm_i8(yq, H, layer["wih_q"], layer["wih_scale"], layer["bias"], b["gx"][i0], CW, rows, 4 * H, H,
                                   rows_inner=TB, valid_inner=TB, stride_inner=1, stride_outer=CS * TB, cb_width=CW,
                                   cb_rows=TB, stream=st, max_ctas=max_ctas)
                return
            with staged("lstm_in_gemm", st):
                native.gemm(src[i0], H, layer["wih"], layer["bias"], b["gx"][i0], CW, cnt * T * TB, 4 * H, H,
                            rows_inner=TB, valid_inner=TB, stride_inner=1, stride_outer=CS * TB, cb_width=CW, cb_rows=TB,
                            impl=gemm_impl, stream=st, max_ctas=max_ctas)

        def rec(dst, layer, tiles, st):
            i0, cnt = tiles
            n = min(cnt * TB, N - i0 * TB)
            with staged("lstm_rec", st):
                native.lstm_rec_tile(b["gx"][i0], layer["whh"], dst[i0], T, n, H, layer["reverse"], stream=st,
                                     workspace=b["hx"][i0])

        def crf_gemm(src, i, st, max_ctas=0):   # rows r = t*TB + i_chunk of tile i -> out[n0 + i_chunk][t]
            n0 = i * TB
            nb = min(TB, N - n0)
            with staged("crf_gemm", st):
                native.gemm(src[i], H, self.wl, self.bl, out[n0:], self.n_scores, T * TB, self.n_scores, H,
                            act=self.act_l, lo=self.lo, hi=self.hi, rows_inner=TB, valid_inner=nb, stride_inner=T,
                            stride_outer=1, impl=gemm_impl, stream=st, max_ctas=max_ctas)

        def gather(buf):            # [tile][T][48][H] -> [T][N][H]
            return buf.permute(1, 0, 2, 3).reshape(T, nt * TB, H)[:, :N].clone()

        if not streams and events is None and not return_features and gemm_impl == native.GEMM_AUTO and not self.quantize \
                and len(self.lstm) <= native.MAX_LSTM_LAYERS and os.environ.get("B200_COARSE_FWD", "1") != "0":
            # the whole encoder from ONE C-ABI call (b200_lstm_crf_fwd); the per-kernel path below runs the same launches
            # one ctypes call at a time (used when per-kernel events or intermediate activations are wanted)
            native.lstm_crf_fwd(self._plan_struct(b, N, L), x, out, stream=main)
            return out

        with _StreamStage("conv_stem", events, main):
            native.conv_stem(x, self.w1, self.b1, self.act1, self.w2, self.b2, self.act2, b["stem"], Lp, self.pad3)
        if return_features:
            feats["stem"] = b["stem"][:N * Lp * self.c2].view(N, Lp, self.c2)[:, self.pad3:self.pad3 + L].clone()
        cur, nxt = b["ya"], b["yb"]

        if not streams:
            # all tiles in one launch: chunk n of the stem -> (tile n / TB, row n % TB); rows r = n*Tp + t
            with staged("conv_gemm", main):
                native.gemm(b["stem"], self.s3 * self.c2, self.w3, self.b3, b["ya"], H, N * Tp, H, self.k3 * self.c2,
                            act=self.act3, rows_inner=Tp, valid_inner=T, stride_inner=TB, stride_outer=1, group=TB,
                            stride_group=T * TB, impl=gemm_impl, stream=main)
            if return_features:
                feats["conv"] = gather(cur)
            for li, layer in enumerate(self.lstm):
                in_gemm(cur, layer, (0, nt), main)
                rec(nxt, layer, (0, nt), main)
                cur, nxt = nxt, cur
                if return_features:
                    feats[f"lstm{li}"] = gather(cur)
            # all tiles in one launch: rows r = (tile*T + t)*TB + i -> out[tile*TB + i][t]; rows of chunks beyond the batch
            # (last tile) land behind the N valid chunks and are cut off by `valid_rows`
            if N % TB == 0:
                with staged("crf_gemm", main):
                    native.gemm(cur, H, self.wl, self.bl, out, self.n_scores, nt * T * TB, self.n_scores, H, act=self.act_l,
                                lo=self.lo, hi=self.hi, rows_inner=TB, valid_inner=TB, stride_inner=T, stride_outer=1,
                                group=T, stride_group=TB * T, impl=gemm_impl, stream=main)
            else:
                full = N // TB
                if full:
                    with staged("crf_gemm", main):
                        native.gemm(cur, H, self.wl, self.bl, out, self.n_scores, full * T * TB, self.n_scores, H,
                                    act=self.act_l, lo=self.lo, hi=self.hi, rows_inner=TB, valid_inner=TB, stride_inner=T,
                                    stride_outer=1, group=T, stride_group=TB * T, impl=gemm_impl, stream=main)
                crf_gemm(cur, nt - 1, main)
            return (out, feats) if return_features else out

        b["start"].record(main)
        for i in range(nt):
            st = b["streams"][i]
            st.wait_event(b["start"])
            conv_gemm(i, st)
        for li, layer in enumerate(self.lstm):
            for i in range(nt):
                st, rs = b["streams"][i], b["rec_streams"][i]
                in_gemm(cur, layer, (i, 1), st, max_ctas=cap if li > 0 else 0)
                b["rec_ready"][i].record(st)
                rs.wait_event(b["rec_ready"][i])
                rec(nxt, layer, (i, 1), rs)
                b["rec_done"][i].record(rs)
                st.wait_event(b["rec_done"][i])
            cur, nxt = nxt, cur
        for i in range(nt):
            st = b["streams"][i]
            crf_gemm(cur, i, st, max_ctas=cap)
            b["done"][i].record(st)
            main.wait_event(b["done"][i])
        return out

    def forward(self, x, out=None, gemm_impl=native.GEMM_AUTO, return_features=False, events=None, tiled=None,
                decode=None, slot=0):
        with torch.cuda.device(self.device):    # streams / events / launches belong to the plan's device, whatever is current
            return self._forward(x, out=out, gemm_impl=gemm_impl, return_features=return_features, events=events,
                                 tiled=tiled, decode=decode, slot=slot)

    def _forward(self, x, out=None, gemm_impl=native.GEMM_AUTO, return_features=False, events=None, tiled=None,
                 decode=None, slot=0):
        """
        x: [N, 1, L] (or [N, L]) fp16 CUDA -> scores [N, T, C] fp16 (no blank column).
        `events`: optional list; (stage, start, end) CUDA events are appended per kernel.
        `tiled`: run the tile-pipelined schedule (default: whenever the batch has more than one 32-chunk tile).
        `decode`: see `forward_tiled` (ignored by the single-stream schedule).
        """
        import os
        if self.tile and os.environ.get("B200_LSTM_TILE", "1") != "0":
            if tiled is None:
                # Default: the layer-by-layer schedule (14 launches per batch).  With two batches in flight on two streams
                # (score_batches, bench.py) it measured faster than per-tile streams (17.8 vs 18.8 ms per 512-chunk batch):
                # the GEMMs / decode of one batch fill the 82 SMs the other batch's recurrent clusters leave free.
                tiled = (not return_features) and x.shape[0] > self.tile and os.environ.get("B200_TILE_STREAMS", "0") != "0"
            return self.forward_tiles(x, out=out, gemm_impl=gemm_impl, events=events, return_features=return_features,
                                      streams=tiled, slot=slot)
        if slot != 0:
            raise NotImplementedError("several batches in flight (slot != 0) need the tile-layout path (hidden size 384)")
        if tiled is None:
            tiled = (not return_features) and x.shape[0] > self.TILE
        if tiled:
            return self.forward_tiled(x, out=out, gemm_impl=gemm_impl, events=events, decode=decode)

        def stage(name):
            return _Stage(name, events)

        if x.dim() == 3:
            x = x[:, 0, :]
        x = x.to(device=self.device, dtype=torch.float16).contiguous()
        N, L = x.shape
        H = self.hidden
        b = self._buffers(N, L)
        T, Tp, Lp = b["T"], b["Tp"], b["Lp"]
        feats = {}

        with stage("conv_stem"):
            native.conv_stem(x, self.w1, self.b1, self.act1, self.w2, self.b2, self.act2, b["stem"], Lp, self.pad3)
        if return_features:
            feats["stem"] = b["stem"][:N * Lp * self.c2].view(N, Lp, self.c2)[:, self.pad3:self.pad3 + L].clone()

        # strided conv: rows r = n*Tp + t are windows of k3*c2 elements, s3*c2 apart; out[t][n][:]
        cur, nxt = b["ya"], b["yb"]
        with stage("conv_gemm"):
            native.gemm(b["stem"], self.s3 * self.c2, self.w3, self.b3, cur, H, N * Tp, H, self.k3 * self.c2,
                        act=self.act3, rows_inner=Tp, valid_inner=T, stride_inner=N, stride_outer=1, impl=gemm_impl)
        if return_features:
            feats["conv"] = cur.clone()

        for i, layer in enumerate(self.lstm):
            with stage("lstm_in_gemm"):
                native.gemm(cur, H, layer["wih"], layer["bias"], b["gx"], 4 * H, T * N, 4 * H, H, impl=gemm_impl)
            with stage("lstm_rec"):
                native.lstm_rec(b["gx"], layer["whh"], nxt, T, N, H, layer["reverse"])
            cur, nxt = nxt, cur
            if return_features:
                feats[f"lstm{i}"] = cur.clone()

        if out is None:
            out = torch.empty(N, T, self.n_scores, dtype=torch.float16, device=self.device)
        # rows r = t*N + n -> out[n][t][:]
        with stage("crf_gemm"):
            native.gemm(cur, H, self.wl, self.bl, out, self.n_scores, T * N, self.n_scores, H,
                        act=self.act_l, lo=self.lo, hi=self.hi,
                        rows_inner=N, valid_inner=N, stride_inner=T, stride_outer=1, impl=gemm_impl)
        return (out, feats) if return_features else out


class _DecodeCache:
    """Decode results the tile-pipelined forward produced ahead of the `beam_search` call that will ask for them."""

    def __init__(self):
        self._entry = None

    @staticmethod
    def _version(t):
        try:
            return t._version
        except RuntimeError:  # inference tensors carry no version counter
            return -1

    def put(self, scores, params, outputs):
        self._entry = (scores.data_ptr(), tuple(scores.shape), self._version(scores), params, outputs)

    def take(self, scores, params):
        e, self._entry = self._entry, None
        if e is not None and e[0] == scores.data_ptr() and e[1] == tuple(scores.shape) \
                and e[2] == self._version(scores) and e[3] == params:
            return e[4]
        return None


DECODE_CACHE = _DecodeCache()


class _LazyStreams:
    """Per-tile streams of the tile-pipelined schedules, created on first use with `native.new_stream` (CUDA streams of their
    own).  `torch.cuda.Stream()` hands out streams from a pool of 32 per device round-robin: the 22 per-tile streams of two buffer
    sets, created eagerly, pushed later requests (the slot and copy streams of `score_batches`) onto pool entries already in use
    -- two "different" streams were then one CUDA stream and batches meant to overlap ran back to back (end-to-end step 17 ->
    20-58 ms, depending on how many streams the process had created before)."""

    def __init__(self, device, n):
        self.device, self.items = device, [None] * n

    def __getitem__(self, i):
        if self.items[i] is None:
            self.items[i] = native.new_stream(self.device)
        return self.items[i]

    def __len__(self):
        return len(self.items)


def compile_lstm_crf(encoder, device, quantize=False):
    return LstmCrfPlan(encoder, device, quantize=quantize)


class CrfDecoder:
    """Workspace-caching wrapper around b200_crf_decode."""

    def __init__(self):
        self._ws_by_device = {}     # one workspace per (device, host thread): basecall() decodes on background threads

    def __call__(self, scores, state_len, blank_score=2.0, qscale=1.0, qbias=0.0, events=None, out=None, slot=0, beam=None):
        """`beam=(beam_width, beam_cut)`: run the beam search (after the forward-backward pass) instead of the exact
        posterior-Viterbi trace-back."""
        with torch.cuda.device(scores.device):
            return self._call(scores, state_len, blank_score, qscale, qbias, events, out, slot, beam)

    def _call(self, scores, state_len, blank_score, qscale, qbias, events, out, slot=0, beam=None):
        """-> (moves, sequence, qstring) uint8 [N, T] on the device; `out`: optional uint8 [3, N, T] to write them into."""
        n, t, c = scores.shape
        if c != 4 ** (state_len + 1):
            raise ValueError(f"scores width {c} does not match state_len {state_len}")
        cached = DECODE_CACHE.take(scores, (state_len, float(blank_score), float(qscale), float(qbias)))
        if cached is not None:
            if out is not None:
                for dst, src in zip(out, cached):
                    dst.copy_(src)
                return tuple(out)
            return cached
        scores = scores.to(torch.float16).contiguous()
        need = native.crf_decode_workspace_bytes(n, t, state_len)
        import threading
        key = (scores.device, threading.get_ident(), slot)
        ws = self._ws_by_device.get(key)
        if ws is None or ws.numel() < need:
            ws = self._ws_by_device[key] = torch.empty(need, dtype=torch.uint8, device=scores.device)
        outs = list(out) if out is not None else [torch.empty(n, t, dtype=torch.uint8, device=scores.device) for _ in range(3)]
        with _Stage("crf_decode", events):
            if beam is None:
                native.crf_decode(scores, state_len, blank_score, qscale, qbias, ws, *outs)
            else:
                native.crf_beam_search(scores, state_len, blank_score, beam[0], beam[1], qscale, qbias, ws, *outs)
        return tuple(outs)  # moves, sequence, qstring
