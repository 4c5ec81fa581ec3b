"""Forward-pass plan builder: reference NetDesc structure -> tensors + ops for the C ABI.

Mirrors models/net_desc.py:144-200 (NetDesc.forward) over models/backbone/resnet.py:273-286
(ResNet34 encoder, 7x7 stride-1 stem) with the B200-first restructuring:
  * BN folded, bias/ReLU/residual fused in the conv epilogue;
  * the first decoder stage input  x3 + up2x(conv_map(x4))  (net_desc.py:183-188) is the same
    for every decoder, so it is computed once and the first 256->256 conv of all decoders
    runs as ONE convolution with the output channels concatenated;
  * the output-head tail (1x1 96->C, softmax, channel drop / argmax, centre crop:
    models/run_desc.py:451-491) is one kernel writing the per-patch canvas whose channel
    layout is the one infer/tile.py:116-134 builds from decoder_kwargs.
"""
from collections import OrderedDict

import numpy as np

from . import _lib
from .pack import BlobBuilder, _np, bn_of, fold_bn, pack_conv, pack_stem, strip_module_prefix

# BasicBlock encoders of models/backbone/resnet.py:292-313 (same block type, same decoder filters:
# models/backbone/__init__.py:27-28); resnet50 (Bottleneck) and the other families are outside the path
RESNET_BLOCKS = {"resnet18": [2, 2, 2, 2], "resnet34": [3, 4, 6, 3]}
RESNET34_BLOCKS = RESNET_BLOCKS["resnet34"]
RESNET34_FILTERS = [64, 64, 128, 256, 512]  # models/backbone/__init__.py resnet18 / resnet34 rows

# models/run_desc.py:472-479
HEAD_NAME_MAP = {
    "Gland": "Gland-INST", "Gland#TYPE": "Gland-TYPE", "Lumen": "Lumen-INST",
    "Nuclei": "Nuclei-INST", "Nuclei#TYPE": "Nuclei-TYPE", "Patch-Class": "Patch-Class",
}


def canvas_layout(decoder_kwargs):
    """Channel table of infer/tile.py:116-134. Returns (idx_dict, nr_channels)."""
    nr = 0
    idx = OrderedDict()
    for tissue_name, info in decoder_kwargs.items():
        for chann_type, nr_chans in info.items():
            start = nr
            if chann_type == "INST":
                nr += nr_chans - 1
                idx[tissue_name + "-INST"] = [start, nr]
            elif chann_type == "TYPE":
                nr += 1
                idx[tissue_name.split("#")[0] + "-TYPE"] = [start, nr]
            else:
                nr += 1
                idx[tissue_name] = [start, nr]
    return idx, nr


class PackedModel:
    """Folded + packed weights of one model directory, independent of the batch shape."""

    def __init__(self, state_dict, model_args):
        name = model_args.get("encoder_backbone_name")
        if name not in RESNET_BLOCKS:
            raise ValueError("cerberus_b200 implements the resnet34 / resnet18 encoders only (got %r); the "
                             "other backbones are outside the hot path (SURVEY.md section 2)" % (name,))
        self.backbone = name
        self.blocks = RESNET_BLOCKS[name]
        self.decoder_kwargs = OrderedDict(
            (k, OrderedDict(v)) for k, v in model_args["decoder_kwargs"].items())
        self.considered_tasks = list(model_args["considered_tasks"])
        for t in self.considered_tasks:
            if t not in self.decoder_kwargs:
                raise KeyError("considered task %r is not in decoder_kwargs" % t)
        self.idx_dict, self.canvas_c = canvas_layout(self.decoder_kwargs)
        sd = strip_module_prefix(state_dict)
        blob = BlobBuilder()
        L = self.layers = {}

        def conv_bn(key, wkey, bnkey, bkey=None):
            w = _np(sd[wkey])
            b = _np(sd[bkey]) if bkey is not None else None
            w, b = fold_bn(w, b, bn_of(sd, bnkey))
            L[key] = pack_conv(blob, w, b)

        # encoder (models/backbone/resnet.py)
        w, b = fold_bn(_np(sd["backbone.conv1.weight"]), None, bn_of(sd, "backbone.bn1"))
        L["stem"] = pack_stem(blob, w, b)
        for li, nblocks in enumerate(self.blocks, start=1):
            for bi in range(nblocks):
                p = "backbone.layer%d.%d" % (li, bi)
                conv_bn(p + ".conv1", p + ".conv1.weight", p + ".bn1")
                conv_bn(p + ".conv2", p + ".conv2.weight", p + ".bn2")
                if (p + ".downsample.0.weight") in sd:
                    conv_bn(p + ".downsample", p + ".downsample.0.weight", p + ".downsample.1")
        L["conv_map"] = pack_conv(blob, _np(sd["conv_map.weight"]), None)

        # decoders, in nn.ModuleDict order = decoder_kwargs order filtered by considered tasks
        self.seg_decoders = [d for d in self.decoder_kwargs
                             if d in self.considered_tasks and d != "Patch-Class"]
        self.has_pclass = "Patch-Class" in self.considered_tasks and "Patch-Class" in self.decoder_kwargs
        first_w, first_b = [], []
        for d in self.seg_decoders:
            for blk in range(4):
                for cv in range(2):
                    p = "decoder_head.%s.%d.block.%d" % (d, blk, cv)
                    w, b = fold_bn(_np(sd[p + ".conv.weight"]), _np(sd[p + ".conv.bias"]),
                                   bn_of(sd, p + ".bn"))
                    if blk == 0 and cv == 0:
                        first_w.append(w)
                        first_b.append(b)
                    else:
                        L["dec.%s.%d.%d" % (d, blk, cv)] = pack_conv(blob, w, b)
            heads = self.decoder_kwargs[d]
            if len(heads) != 1:
                raise ValueError("decoder %r: exactly one output head expected" % d)
            (clf, nclass), = heads.items()
            p = "output_head.%s.%s.x" % (d, clf)
            w, b = fold_bn(_np(sd[p + ".0.block.0.conv.weight"]), _np(sd[p + ".0.block.0.conv.bias"]),
                           bn_of(sd, p + ".0.block.0.bn"))
            L["head.%s.hidden" % d] = pack_conv(blob, w, b)
            w2 = _np(sd[p + ".1.conv.weight"]).reshape(nclass, 96)
            b2 = _np(sd[p + ".1.conv.bias"])
            L["head.%s.out" % d] = {"w_off": blob.add(w2.astype(np.float32)),
                                    "b_off": blob.add(b2.astype(np.float32)),
                                    "classes": nclass, "clf": clf}
        if self.seg_decoders:
            L["dec.first"] = pack_conv(blob, np.concatenate(first_w, 0), np.concatenate(first_b, 0))
        if self.has_pclass:
            p = "decoder_head.Patch-Class"
            bn1 = bn_of(sd, p + ".bn1")
            s1 = bn1["weight"] / np.sqrt(bn1["running_var"] + 1e-5)
            sh1 = bn1["bias"] - bn1["running_mean"] * s1
            w1, b1 = fold_bn(_np(sd[p + ".conv1.weight"]), _np(sd[p + ".conv1.bias"]), bn_of(sd, p + ".bn2"))
            w2 = _np(sd[p + ".conv2.weight"])
            b2 = _np(sd[p + ".conv2.bias"])
            ncls = w2.shape[0]
            params = np.concatenate([s1, sh1, w1.reshape(256, 512).T.ravel(), b1,
                                     w2.reshape(ncls, 256).ravel(), b2]).astype(np.float32)
            L["pclass"] = {"w_off": blob.add(params), "classes": ncls}
        self.blob = blob.finish()


def c_model_tables(model):
    """PackedModel -> (cerb_model_desc, cerb_layer array) for cerb_model_create / cerb_model_spec:
    the layer table of the blob in the role / index vocabulary of include/cerberus_b200.h."""
    L = model.layers
    rows = []

    def add(role, layer, a=0, b=0, c=0, classes=0):
        rows.append(_lib.Layer(role, a, b, c, layer.get("cout", 0), layer.get("cin", 0),
                               layer.get("kh", 0), layer.get("kw", 0), layer.get("w_shift", 0),
                               classes, layer["w_off"], layer.get("w_lo_off", -1),
                               layer.get("b_off", -1)))

    add(_lib.L_STEM, L["stem"])
    for li, nblocks in enumerate(getattr(model, "blocks", RESNET34_BLOCKS), start=1):
        for bi in range(nblocks):
            p = "backbone.layer%d.%d" % (li, bi)
            add(_lib.L_BLOCK_CONV1, L[p + ".conv1"], li, bi)
            add(_lib.L_BLOCK_CONV2, L[p + ".conv2"], li, bi)
            if (p + ".downsample") in L:
                add(_lib.L_BLOCK_DOWN, L[p + ".downsample"], li, bi)
    add(_lib.L_CONV_MAP, L["conv_map"])
    desc = _lib.ModelDesc()
    desc.n_decoders = len(model.seg_decoders)
    if desc.n_decoders > _lib.MAX_DECODERS:
        raise ValueError("more than %d segmentation decoders" % _lib.MAX_DECODERS)
    if model.seg_decoders:
        add(_lib.L_DEC_FIRST, L["dec.first"])
    for di, d in enumerate(model.seg_decoders):
        for blk in range(4):
            for cv in range(2):
                if blk == 0 and cv == 0:
                    continue
                add(_lib.L_DEC_CONV, L["dec.%s.%d.%d" % (d, blk, cv)], di, blk, cv)
        add(_lib.L_HEAD_HIDDEN, L["head.%s.hidden" % d], di)
        ho = L["head.%s.out" % d]
        add(_lib.L_HEAD_OUT, ho, di, classes=ho["classes"])
        desc.head_mode[di] = _lib.HEAD_INST if ho["clf"] == "INST" else _lib.HEAD_TYPE
        desc.classes[di] = ho["classes"]
        desc.canvas_coff[di] = model.idx_dict[HEAD_NAME_MAP[d]][0]
    desc.has_pclass = int(model.has_pclass)
    if model.has_pclass:
        add(_lib.L_PCLASS, L["pclass"], classes=L["pclass"]["classes"])
        desc.pclass_classes = L["pclass"]["classes"]
        desc.pclass_canvas_coff = model.idx_dict["Patch-Class"][0]
    desc.canvas_c = model.canvas_c
    arr = (_lib.Layer * len(rows))(*rows)
    return desc, arr


class PlanSpec:
    """Tensors + ops for one batch shape. Pure host data (testable without a GPU)."""

    def __init__(self, model, n, h, w, out_h, out_w, want_logits=False, fuse_head=True,
                 fuse_upadd=False, fuse_tail=False, level_sync=False):
        if h % 16 or w % 16:
            raise ValueError("input size must be a multiple of 16 (got %dx%d)" % (h, w))
        if out_h > h or out_w > w:
            raise ValueError("output shape exceeds the input shape")
        self.model = model
        self.n, self.h, self.w, self.out_h, self.out_w = n, h, w, out_h, out_w
        self.tensors = []
        self.ops = []
        self.logit_tensors = OrderedDict()
        self.named = {}
        L = model.layers
        T = self._tensor
        F = RESNET34_FILTERS

        t_in = T("input", n, h, w, 3, _lib.CERB_U8)
        t_prep = T("prep", n, h, w + 8, 8)
        self._op(_lib.OP_PREP, in0=t_in, out=t_prep)
        x0 = T("x0", n, h, w, 64)
        self._conv(L["stem"], t_prep, x0, relu=1, stem=1)
        hs = [h, h // 2, h // 4, h // 8, h // 16]
        ws = [w, w // 2, w // 4, w // 8, w // 16]
        pool = T("pool", n, hs[1], ws[1], 64)
        self._op(_lib.OP_MAXPOOL, in0=x0, out=pool)
        feats = [x0]
        cur = pool
        for li, nblocks in enumerate(getattr(model, "blocks", RESNET34_BLOCKS), start=1):
            c = F[li]
            mid = T("l%d.mid" % li, n, hs[li], ws[li], c)
            o = T("l%d.o" % li, n, hs[li], ws[li], c)
            ds = T("l%d.ds" % li, n, hs[li], ws[li], c) if li > 1 else None
            pair = (pool, o) if li == 1 else (ds, o)
            for bi in range(nblocks):
                p = "backbone.layer%d.%d" % (li, bi)
                stride = 2 if (li > 1 and bi == 0) else 1
                self._conv(L[p + ".conv1"], cur, mid, relu=1, stride=stride)
                if (p + ".downsample") in L:
                    self._conv(L[p + ".downsample"], cur, ds, relu=0, stride=2)
                    res, dst = ds, o
                else:
                    # never in place: the residual (= block input) is read by the epilogue
                    res = cur
                    dst = pair[0] if cur == pair[1] else pair[1]
                self._conv(L[p + ".conv2"], mid, dst, relu=1, residual=res)
                cur = dst
            self.named["x%d" % li] = cur
            feats.append(cur)
        x0, x1, x2, x3, x4 = feats
        self.named["x0"] = x0

        canvas = T("canvas", n, out_h, out_w, model.canvas_c, _lib.CERB_F32)
        self.canvas = canvas
        if model.has_pclass:
            # depends on x4 only and nothing in the plan reads its canvas channel: queued right
            # after the encoder and flagged `side`, it overlaps the decoders in the replayed graph
            pc = L["pclass"]
            lg = -1
            if want_logits:
                lg = T("logits.Patch-Class", n, 1, 1, pc["classes"], _lib.CERB_F32)
            self._op(_lib.OP_PCLASS, in0=x4, out=canvas, out_coff=model.idx_dict["Patch-Class"][0],
                     cout=pc["classes"], logits_out=lg, w_off=pc["w_off"], side=1)
            pclass_logits = lg
        D = len(model.seg_decoders)
        if D:
            f4 = T("conv_map", n, hs[4], ws[4], 256)
            self._conv(L["conv_map"], x4, f4, relu=0)
            s3 = T("s3", n, hs[3], ws[3], 256)
            self._op(_lib.OP_UPADD, in0=x3, in1=f4, out=s3)
            u4a = T("u4a", n, hs[3], ws[3], 256 * D)
            self._conv(L["dec.first"], s3, u4a, relu=1)
            if level_sync and fuse_head and not (fuse_upadd or fuse_tail):
                self._decoders_level_sync(model, L, T, n, h, w, hs, ws, feats, u4a, canvas, want_logits)
                D = 0  # handled
        if D:
            u4b = T("u4b", n, hs[3], ws[3], 128)
            s2 = T("s2", n, hs[2], ws[2], 128)
            a2 = T("a2", n, hs[2], ws[2], 128)
            b2 = T("b2", n, hs[2], ws[2], 64)
            s1 = T("s1", n, hs[1], ws[1], 64) if not fuse_upadd else -1
            a1 = T("a1", n, hs[1], ws[1], 64)
            b1 = T("b1", n, hs[1], ws[1], 64)
            s0 = T("s0", n, h, w, 64) if not fuse_upadd else -1
            a0 = T("a0", n, h, w, 64)
            b0 = T("b0", n, h, w, 64) if not fuse_tail else -1
            hid = T("hid", n, h, w, 96) if not (fuse_head or fuse_tail) else -1
            for di, d in enumerate(model.seg_decoders):
                self._conv(L["dec.%s.0.1" % d], u4a, u4b, relu=1, in_coff=di * 256)
                self._op(_lib.OP_UPADD, in0=x2, in1=u4b, out=s2)
                self._conv(L["dec.%s.1.0" % d], s2, a2, relu=1)
                self._conv(L["dec.%s.1.1" % d], a2, b2, relu=1)
                if fuse_upadd:  # s1 = x1 + up(b2) is built inside the conv's producer
                    self._conv(L["dec.%s.2.0" % d], x1, a1, relu=1, up_prev1=b2 + 1)
                else:
                    self._op(_lib.OP_UPADD, in0=x1, in1=b2, out=s1)
                    self._conv(L["dec.%s.2.0" % d], s1, a1, relu=1)
                self._conv(L["dec.%s.2.1" % d], a1, b1, relu=1)
                if fuse_upadd:
                    self._conv(L["dec.%s.3.0" % d], x0, a0, relu=1, up_prev1=b1 + 1)
                else:
                    self._op(_lib.OP_UPADD, in0=x0, in1=b1, out=s0)
                    self._conv(L["dec.%s.3.0" % d], s0, a0, relu=1)
                ho = L["head.%s.out" % d]
                key = HEAD_NAME_MAP[d]
                lo, hi_ = model.idx_dict[key]
                lg = -1
                if want_logits:
                    lg = T("logits." + key, n, h, w, ho["classes"], _lib.CERB_F32)
                    self.logit_tensors[key] = lg
                mode = _lib.HEAD_INST if ho["clf"] == "INST" else _lib.HEAD_TYPE
                if fuse_tail:
                    # last decoder conv + the whole output head in ONE kernel (fp16 mode): neither
                    # the 64-channel decoder output nor the 96-channel hidden tensor touches HBM
                    hd = L["head.%s.hidden" % d]
                    self._conv(L["dec.%s.3.1" % d], a0, canvas, relu=1, out_coff=lo,
                               aux_classes=ho["classes"], aux_w_off=ho["w_off"],
                               aux_b_off=ho["b_off"], head_mode=mode, logits_out=lg,
                               tail_w_off=hd["w_off"], tail_b_off=hd["b_off"],
                               tail_w_shift=hd.get("w_shift", 0))
                    continue
                self._conv(L["dec.%s.3.1" % d], a0, b0, relu=1)
                if fuse_head:
                    # 1x1 64->96 + BN + ReLU + 1x1 96->C + softmax/argmax/crop in ONE kernel: the
                    # 96-channel hidden tensor never exists in HBM
                    self._conv(L["head.%s.hidden" % d], b0, canvas, relu=1, out_coff=lo,
                               aux_classes=ho["classes"], aux_w_off=ho["w_off"],
                               aux_b_off=ho["b_off"], head_mode=mode, logits_out=lg)
                else:
                    self._conv(L["head.%s.hidden" % d], b0, hid, relu=1)
                    self._op(_lib.OP_HEAD, in0=hid, out=canvas, out_coff=lo, cout=ho["classes"],
                             head_mode=mode, logits_out=lg, w_off=ho["w_off"], b_off=ho["b_off"])
        if model.has_pclass and want_logits:
            self.logit_tensors["Patch-Class"] = pclass_logits

    def _decoders_level_sync(self, model, L, T, n, h, w, hs, ws, feats, u4a, canvas, want_logits):
        """OPT-IN (level_sync=True / CERB_LEVEL_SYNC=1), measured and NOT faster on B200: with
        five separate passes `upadd` already runs at 5.9 TB/s (0.102 ms per decoder at 256^2);
        the grouped pass moves 36 % fewer bytes but is write-dominated (1.34 of 1.94 GB) and takes
        0.545 ms for the five, and the step went from 6.12 to 6.21 ms (profiles/r2_level_sync_ab.txt).
        The D decoders level by level instead of decoder by decoder: the skip tensor of a level
        (x2 / x1 / x0) is the same for every decoder (models/net_desc.py:183-188), so ONE grouped
        UPADD reads it once and writes the D sums (at 256^2, batch 32: 1.9 GB of traffic instead
        of 3.0 GB); consecutive convolutions of a level are independent of each other, so each
        one's prologue overlaps its predecessor's tail. Every decoder owns its tensors (ids of a
        kind are consecutive: the grouped UPADD addresses them as first id + d)."""
        x0, x1, x2 = feats[0], feats[1], feats[2]
        D = len(model.seg_decoders)
        decs = model.seg_decoders

        def per_decoder(name, hh, ww, c):
            return [T("%s.%d" % (name, d), n, hh, ww, c) for d in range(D)]

        u4b = per_decoder("u4b", hs[3], ws[3], 128)
        s2 = per_decoder("s2", hs[2], ws[2], 128)
        a2 = per_decoder("a2", hs[2], ws[2], 128)
        b2 = per_decoder("b2", hs[2], ws[2], 64)
        s1 = per_decoder("s1", hs[1], ws[1], 64)
        a1 = per_decoder("a1", hs[1], ws[1], 64)
        b1 = per_decoder("b1", hs[1], ws[1], 64)
        s0 = per_decoder("s0", h, w, 64)
        a0 = per_decoder("a0", h, w, 64)
        b0 = per_decoder("b0", h, w, 64)

        def upadd(skip, prev, out):
            if D > 1:
                self._op(_lib.OP_UPADD, in0=skip, in1=prev[0], out=out[0], cout=D)
            else:
                self._op(_lib.OP_UPADD, in0=skip, in1=prev[0], out=out[0])

        for di, d in enumerate(decs):
            self._conv(L["dec.%s.0.1" % d], u4a, u4b[di], relu=1, in_coff=di * 256)
        upadd(x2, u4b, s2)
        for di, d in enumerate(decs):
            self._conv(L["dec.%s.1.0" % d], s2[di], a2[di], relu=1)
        for di, d in enumerate(decs):
            self._conv(L["dec.%s.1.1" % d], a2[di], b2[di], relu=1)
        upadd(x1, b2, s1)
        for di, d in enumerate(decs):
            self._conv(L["dec.%s.2.0" % d], s1[di], a1[di], relu=1)
        for di, d in enumerate(decs):
            self._conv(L["dec.%s.2.1" % d], a1[di], b1[di], relu=1)
        upadd(x0, b1, s0)
        for di, d in enumerate(decs):
            self._conv(L["dec.%s.3.0" % d], s0[di], a0[di], relu=1)
        for di, d in enumerate(decs):
            self._conv(L["dec.%s.3.1" % d], a0[di], b0[di], relu=1)
        for di, d in enumerate(decs):
            ho = L["head.%s.out" % d]
            key = HEAD_NAME_MAP[d]
            lo, _ = model.idx_dict[key]
            lg = -1
            if want_logits:
                lg = T("logits." + key, n, h, w, ho["classes"], _lib.CERB_F32)
                self.logit_tensors[key] = lg
            mode = _lib.HEAD_INST if ho["clf"] == "INST" else _lib.HEAD_TYPE
            # 1x1 64->96 + BN + ReLU + 1x1 96->C + softmax/argmax/crop in ONE kernel
            self._conv(L["head.%s.hidden" % d], b0[di], canvas, relu=1, out_coff=lo,
                       aux_classes=ho["classes"], aux_w_off=ho["w_off"],
                       aux_b_off=ho["b_off"], head_mode=mode, logits_out=lg)

    # -- helpers
    def _tensor(self, name, n, h, w, c, dtype=_lib.CERB_F16):
        self.tensors.append((name, n, h, w, c, dtype))
        self.named[name] = len(self.tensors) - 1
        return len(self.tensors) - 1

    def _op(self, kind, **kw):
        d = dict(kind=kind, in0=-1, in1=-1, out=-1, in_coff=0, in_c=0, out_coff=0, cout=0, kh=0,
                 kw=0, stride=0, pad=0, relu=0, stem=0, head_mode=0, logits_out=-1, w_off=-1,
                 w_lo_off=-1, b_off=-1, box_w=0, w_shift=0, aux_classes=0, aux_w_off=-1,
                 aux_b_off=-1, up_prev1=0, tail_w_off=-1, tail_b_off=-1, tail_w_shift=0, side=0)
        d.update(kw)
        self.ops.append(d)

    def _conv(self, layer, src, dst, relu, stride=1, residual=-1, stem=0, in_coff=0, out_coff=0,
              **extra):
        k = layer["kh"]
        self._op(_lib.OP_CONV, in0=src, in1=residual, out=dst, in_coff=in_coff,
                 in_c=(8 if stem else layer["cin"]), out_coff=out_coff, cout=layer["cout"], kh=k,
                 kw=k, stride=stride, pad=k // 2, relu=relu, stem=stem, w_off=layer["w_off"],
                 w_lo_off=layer["w_lo_off"], b_off=layer["b_off"], w_shift=layer.get("w_shift", 0),
                 **extra)

    def conv_flops(self):
        """Algorithmic conv FLOPs (2*M*N*K, no padding / zero-weight credit) of one batch."""
        total = 0
        for op in self.ops:
            if op["kind"] == _lib.OP_CONV:
                _, n, h, w, _, _ = self.tensors[op["in0"] if op["aux_classes"] else op["out"]]
                cin = 3 if op["stem"] else op["in_c"]
                total += 2 * n * h * w * op["cout"] * op["kh"] * op["kw"] * cin
                total += 2 * n * h * w * op["aux_classes"] * 96
                if op["tail_w_off"] >= 0:
                    total += 2 * n * h * w * 96 * 64
            elif op["kind"] == _lib.OP_HEAD:
                _, n, h, w, _, _ = self.tensors[op["in0"]]
                total += 2 * n * h * w * op["cout"] * 96
            elif op["kind"] == _lib.OP_PCLASS:
                total += 2 * self.n * (512 * 256 + 256 * op["cout"])
        return total

    def c_arrays(self):
        td = (_lib.TensorDesc * len(self.tensors))()
        for i, (_, n, h, w, c, dt) in enumerate(self.tensors):
            td[i] = _lib.TensorDesc(n, h, w, c, dt)
        ops = (_lib.Op * len(self.ops))()
        for i, d in enumerate(self.ops):
            o = _lib.Op()
            for k, v in d.items():
                setattr(o, k, v)
            ops[i] = o
        return td, ops
