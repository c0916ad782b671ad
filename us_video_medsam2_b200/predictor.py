"""Drop-in mirror of the reference's video predictor API on the B200 kernel path.

Same class / method names, argument meaning, return types, error behaviour and `inference_state`
keys as sam2/sam2_video_predictor.py:18-1172 (and the NPZ variant sam2_video_predictor_npz.py:44-115),
so a driver written against the reference (`init_state` / `add_new_mask` / `add_new_points_or_box` /
`propagate_in_video` / `reset_state` / ...) runs unchanged.  The module holds the reference's
state-dict ABI (471 tensors, state_dict_abi.json) as ordinary nn.Parameters, so `load_state_dict`,
strict checkpoint loading, `.to(device)` and `.eval()` behave as on the reference; the compute is
done by `engine.Engine` on weights re-packed for the kernels (re-packed lazily whenever the
parameters change).  There is no CPU path: the model must live on a CUDA device.
"""
import os
import warnings
from collections import OrderedDict

import torch
from torch import nn

from . import _lib, ops, synth
from .engine import NO_OBJ_SCORE, Engine, EtamTiConfig, HieraBPlusConfig, ModelConfig, PackedWeights

try:  # progress bar like the reference (sam2_video_predictor.py:703); optional
    from tqdm import tqdm as _tqdm
except Exception:  # pragma: no cover
    _tqdm = None


def _progress(it, desc):
    if _tqdm is None or os.environ.get("TQDM_DISABLE") == "1":
        return it
    return _tqdm(it, desc=desc)


class _Node(nn.Module):
    """Bare container used to rebuild the reference's module tree (names only, no forward)."""


def _install_abi_parameters(root, abi=None):
    for name, shape in (abi if abi is not None else synth.state_dict_abi()):
        *path, leaf = name.split(".")
        mod = root
        for p in path:
            if p not in mod._modules:
                mod.add_module(p, _Node())
            mod = mod._modules[p]
        t = torch.zeros(shape, dtype=torch.float32)
        if leaf == "positional_encoding_gaussian_matrix":
            mod.register_buffer(leaf, t)
        else:
            mod.register_parameter(leaf, nn.Parameter(t, requires_grad=False))


class SAM2VideoPredictor(nn.Module):
    """The predictor class to handle user interactions and manage inference states."""

    _config_base = ModelConfig           # architecture constants of the shipped configuration
    _abi = staticmethod(synth.state_dict_abi)   # [(parameter name, shape)] of the reference's state dict
    # other architectures the same class serves (the `variant` key of their YAML): Hiera-B+ at 1024^2
    _variants = {"hiera_b+": (HieraBPlusConfig, synth.bplus_state_dict_abi)}

    def __init__(self, fill_hole_area=0, non_overlap_masks=False, clear_non_cond_mem_around_input=False,
                 clear_non_cond_mem_for_multi_obj=False, add_all_frames_to_correct_as_cond=False,
                 encoder_batch=1, use_cuda_graphs=True, encoder_sms=0, variant=None, **model_kwargs):
        super().__init__()
        if variant is not None:
            self._config_base, self._abi = self._variants[variant]
        cfg = type("Cfg", (self._config_base,), {})
        for k, v in model_kwargs.items():
            if k in ("image_encoder", "memory_attention", "memory_encoder", "sam_mask_decoder_extra_args",
                     "compile_image_encoder"):
                continue  # architecture blocks are validated by build_sam; decoder args handled below
            if hasattr(cfg, k):
                setattr(cfg, k, v)
        extra = model_kwargs.get("sam_mask_decoder_extra_args") or {}
        cfg.dynamic_multimask_via_stability = bool(extra.get("dynamic_multimask_via_stability", False))
        cfg.dynamic_multimask_stability_delta = float(extra.get("dynamic_multimask_stability_delta", 0.05))
        cfg.dynamic_multimask_stability_thresh = float(extra.get("dynamic_multimask_stability_thresh", 0.98))
        cfg.fill_hole_area = fill_hole_area
        # class default of SAM2Base (sam2_base.py:60-62) unless the builder's post-processing override passes it
        cfg.binarize_mask_from_pts_for_mem_enc = bool(model_kwargs.get("binarize_mask_from_pts_for_mem_enc", False))
        self.cfg = cfg
        self.image_size = cfg.image_size
        self.num_maskmem = cfg.num_maskmem
        self.hidden_dim = cfg.d_model
        self.mem_dim = cfg.mem_dim
        self.fill_hole_area = fill_hole_area
        self.non_overlap_masks = non_overlap_masks
        self.non_overlap_masks_for_mem_enc = cfg.non_overlap_masks_for_mem_enc
        self.clear_non_cond_mem_around_input = clear_non_cond_mem_around_input
        self.clear_non_cond_mem_for_multi_obj = clear_non_cond_mem_for_multi_obj
        self.add_all_frames_to_correct_as_cond = add_all_frames_to_correct_as_cond
        self.memory_temporal_stride_for_eval = cfg.memory_temporal_stride_for_eval
        # frames encoded per image-encoder launch group (frame-parallel encoder); 1 = reference order
        self.encoder_batch = max(1, int(encoder_batch))
        # steady-state tracked frames are replayed from a CUDA graph keyed by (objects, #memories, #pointers, H, W)
        self.use_cuda_graphs = bool(use_cuda_graphs)
        # > 0: during propagate_in_video the look-ahead encoder batches run CONCURRENTLY with the tracked frames, on a
        # green-context stream that owns this many SMs (pipeline.SmPartition); 0 = encoder and tracking alternate
        self.encoder_sms = int(os.environ.get("USVM2_ENCODER_SMS", encoder_sms))
        self._partition_obj, self._partition_error, self._remote = None, None, None
        self._static_gen = 0            # bumped by every replay of the full-device encoder graph (static output buffers)
        self._pipeline_owner = None     # the session whose propagate_in_video currently owns the look-ahead slots
        _install_abi_parameters(self, self._abi())
        self._engine = None
        self._engine_key = None
        self._graphs, self._graph_seen, self._ctrl = {}, {}, None
        # steady-state frames are software-pipelined (USVM2_PIPELINE_FRAMES=0 or the attribute turns it off): the graph of
        # frame t also computes the part of frame t + 1's memory attention that does not depend on frame t's memory
        # (Engine.attention_prefix), beside the decoder / memory encoder.  Bit-identical results; +1.5 % frames/s
        # (DESIGN.md section 10).
        self.pipeline_frames = os.environ.get("USVM2_PIPELINE_FRAMES", "1") != "0"
        self._pipe_state, self._pipe_bufs, self._last_frame_key = None, {}, None

    # ------------------------------------------------------------------ module plumbing
    @property
    def device(self):
        return next(self.parameters()).device

    def forward(self, *args, **kwargs):
        raise NotImplementedError("Please use the corresponding methods in SAM2VideoPredictor for inference")

    def _weights_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters()) + (str(self.device),)

    def _sync_engine(self):
        """(Re)pack the kernel weights if the parameters changed (load_state_dict / .to()); called once per
        public API call so the per-frame loop never pays for the check."""
        if self.device.type != "cuda":
            raise RuntimeError("us_video_medsam2_b200 has no CPU path: move the predictor to a CUDA device "
                               "(the kernels target sm_100a)")
        _lib.bind_device(self.device.index if self.device.index is not None else torch.cuda.current_device())
        key = self._weights_key()
        if self._engine is None or key != self._engine_key:
            sd = {k: v for k, v in self.state_dict().items()}
            self._engine = Engine(PackedWeights(sd, self.device, self.cfg))
            if self._partition_obj is not None:
                self._engine.sm_budget = self._partition_obj.total_sms - self._partition_obj.sms
            self._engine_key = key
            self._graphs, self._graph_seen = {}, {}  # captured graphs bake the old weight pointers
            self._pipe_state, self._pipe_bufs = None, {}
            self._ctrl = ops.new_frame_ctrl(self.device)
        return self._engine

    def engine(self):
        return self._engine if self._engine is not None else self._sync_engine()

    @classmethod
    def from_pretrained(cls, model_id, **kwargs):
        raise RuntimeError("from_pretrained needs network access to the Hugging Face hub, which this build does "
                           "not use; build with build_sam2_video_predictor(config, ckpt_path) instead")

    # ------------------------------------------------------------------ session state
    @torch.inference_mode()
    def init_state(self, video_path, offload_video_to_cpu=False, offload_state_to_cpu=False,
                   async_loading_frames=False):
        """Initialize an inference state from a JPEG folder (reference :44-111)."""
        from .frames import load_video_frames

        self._sync_engine()
        images, vh, vw = load_video_frames(video_path, self.image_size, offload_video_to_cpu,
                                           async_loading_frames=async_loading_frames, compute_device=self.device)
        return self._new_state(images, vh, vw, offload_video_to_cpu, offload_state_to_cpu)

    def _new_state(self, images, video_height, video_width, offload_video_to_cpu, offload_state_to_cpu):
        if offload_state_to_cpu:
            # the reference moves per-frame outputs to host memory to fit long videos on small GPUs
            # (sam2_video_predictor.py:70-75); here the whole session state of a 512-frame, 4-object clip is < 400 MB
            # of the 180 GB of HBM, so the flag is accepted and the state simply stays resident
            warnings.warn("offload_state_to_cpu is accepted but has no effect: the session state stays in HBM")
        self._sync_engine()
        dev = self.device
        st = {}
        st["images"] = images
        st["num_frames"] = len(images)
        st["offload_video_to_cpu"] = offload_video_to_cpu
        st["offload_state_to_cpu"] = offload_state_to_cpu
        st["video_height"] = video_height
        st["video_width"] = video_width
        st["device"] = dev
        st["storage_device"] = dev
        st["point_inputs_per_obj"] = {}
        st["mask_inputs_per_obj"] = {}
        st["cached_features"] = {}
        st["constants"] = {}
        st["obj_id_to_idx"] = OrderedDict()
        st["obj_idx_to_id"] = OrderedDict()
        st["obj_ids"] = []
        st["output_dict"] = {"cond_frame_outputs": {}, "non_cond_frame_outputs": {}}
        st["output_dict_per_obj"] = {}
        st["temp_output_dict_per_obj"] = {}
        st["consolidated_frame_inds"] = {"cond_frame_outputs": set(), "non_cond_frame_outputs": set()}
        st["tracking_has_started"] = False
        st["frames_already_tracked"] = {}
        self._get_image_feature(st, 0)  # warm up + cache frame 0 like the reference (:110)
        return st

    def _obj_id_to_idx(self, st, obj_id):
        idx = st["obj_id_to_idx"].get(obj_id, None)
        if idx is not None:
            return idx
        if st["tracking_has_started"]:
            raise RuntimeError(f"Cannot add new object id {obj_id} after tracking starts. "
                               f"All existing object ids: {st['obj_ids']}. "
                               f"Please call 'reset_state' to restart from scratch.")
        idx = len(st["obj_id_to_idx"])
        st["obj_id_to_idx"][obj_id] = idx
        st["obj_idx_to_id"][idx] = obj_id
        st["obj_ids"] = list(st["obj_id_to_idx"])
        st["point_inputs_per_obj"][idx] = {}
        st["mask_inputs_per_obj"][idx] = {}
        st["output_dict_per_obj"][idx] = {"cond_frame_outputs": {}, "non_cond_frame_outputs": {}}
        st["temp_output_dict_per_obj"][idx] = {"cond_frame_outputs": {}, "non_cond_frame_outputs": {}}
        return idx

    def _tracked_info(self, st, obj_idx, frame_idx):
        """{"reverse": bool} if the frame was already tracked (for this object), else None (reference :238-247)."""
        return st["frames_already_tracked"].get(frame_idx)

    def _obj_idx_to_id(self, st, obj_idx):
        return st["obj_idx_to_id"][obj_idx]

    def _get_obj_num(self, st):
        return len(st["obj_idx_to_id"])

    # ------------------------------------------------------------------ frame store
    def _ensure_store(self, st):
        """Device store of per-frame results (memory, pointer, score, masks), indexed by frame number."""
        B = self._get_obj_num(st)
        store = st.get("_store")
        if store is not None and store.num_frames == st["num_frames"] and 0 < store.B < B and st["tracking_has_started"]:
            # objects added after tracking started (the EfficientTAM predictor allows it): keep the existing objects'
            # results, re-point every stored entry at the grown store
            store.grow(B)
            self._refresh_views(st)
        elif store is None or store.B != B or store.num_frames != st["num_frames"]:
            store = ops.FrameStore(st["num_frames"], B, self.device, T=self.cfg.feat ** 2, low=self.image_size // 4)
            st["_store"] = store
        return store

    def _refresh_views(self, st):
        """Rebuild every stored output entry as views of the (re-laid-out) frame store."""
        for storage_key in ("cond_frame_outputs", "non_cond_frame_outputs"):
            for frame_idx, old in list(st["output_dict"][storage_key].items()):
                st["output_dict"][storage_key][frame_idx] = self._slot_views(st, frame_idx,
                                                                             old["maskmem_features"] is not None)
            for obj_idx, od in st["output_dict_per_obj"].items():
                for frame_idx, old in list(od[storage_key].items()):
                    od[storage_key][frame_idx] = self._obj_slot_views(st, frame_idx, obj_idx,
                                                                      old["maskmem_features"] is not None)

    def _obj_slot_views(self, st, t, obj_idx, with_memory=True):
        """One object's entry of frame t as views into the frame store (`output_dict_per_obj`, reference :747-774)."""
        store = st["_store"]
        s = slice(obj_idx, obj_idx + 1)
        out = {"maskmem_features": None, "maskmem_pos_enc": None, "_mem_tok": None, "_slot": t,
               "pred_masks": store.masks[t][s], "obj_ptr": store.ptr[t][s], "object_score_logits": store.score[t][s]}
        if with_memory:
            out["_mem_tok"] = store.mem[t][s]
            out["maskmem_features"] = self._mem_view(store.mem[t][s])
            out["maskmem_pos_enc"] = self._maskmem_pos_enc(st, 1)
        return out

    def _slot_views(self, st, t, with_memory=True):
        """The `inference_state` entry of frame t as views into the frame store (reference keys :971-977)."""
        store = st["_store"]
        B = store.B
        out = {"maskmem_features": None, "maskmem_pos_enc": None, "_mem_tok": None, "_slot": t,
               "pred_masks": store.masks[t], "obj_ptr": store.ptr[t], "object_score_logits": store.score[t]}
        if with_memory:
            out["_mem_tok"] = store.mem[t]
            out["maskmem_features"] = self._mem_view(store.mem[t])
            out["maskmem_pos_enc"] = self._maskmem_pos_enc(st, B)
        return out

    def _commit(self, st, t, out):
        """Copy a consolidated output (fresh tensors) into slot t of the store and return the store-backed entry."""
        store = self._ensure_store(st)
        store.mem[t].copy_(out["_mem_tok"])
        store.ptr[t].copy_(out["obj_ptr"])
        store.score[t].copy_(out["object_score_logits"])
        store.masks[t].copy_(out["pred_masks"])
        return self._slot_views(st, t)

    # ------------------------------------------------------------------ image features
    def _get_image_feature(self, st, frame_idx, lookahead=None):
        """Per-frame backbone features (reference :879-910).  On a miss, `encoder_batch` consecutive frames in
        the tracking direction are encoded in one batched pass (the image encoder is frame-independent); during
        propagate_in_video a FeaturePipeline may already hold them (encoded ahead on an SM partition or another GPU)."""
        cache = st["cached_features"]
        hit = cache.get(frame_idx)
        if hit is not None and hit.get("_static") and hit.get("_gen") != self._static_gen:
            # views of the encoder graph's static outputs that another session (or a later batch) has overwritten since
            st["cached_features"] = cache = {t: v for t, v in cache.items() if not v.get("_static")}
            hit = None
        if hit is not None:
            return hit
        pipe = st.get("_pipeline")
        if pipe is not None and lookahead is not None:
            got = pipe.get(frame_idx)
            if got is not None:
                return got
        eng = self.engine()
        step = lookahead if lookahead is not None else 0
        n = self.encoder_batch if step != 0 else 1
        idxs = [frame_idx + i * step for i in range(n)] if step != 0 else [frame_idx]
        idxs = [t for t in idxs if 0 <= t < st["num_frames"]]
        if self.use_cuda_graphs and len(idxs) == self.encoder_batch and self.encoder_batch > 1:
            # full look-ahead batch: replay the captured image-encoder graph.  Its outputs are static buffers that the
            # next replay overwrites, so the cache holds exactly the frames of the latest batch.
            graph, static_in, out, n_kernels = self._encoder_graph(len(idxs))
            self._load_frames(st, idxs, static_in)
            graph.replay()
            self._static_gen += 1
            _lib.launch_count += n_kernels
            keep = {}
        else:
            imgs = torch.stack([self._frame(st, t) for t in idxs]).contiguous()
            out = eng.encode_frames(imgs)
            keep = {t: v for t, v in cache.items() if not v.get("_static")} if len(cache) < 4 * self.encoder_batch else {}
        for j, t in enumerate(idxs):
            keep[t] = {k: v[j] for k, v in out.items()}
        if self.use_cuda_graphs and len(idxs) == self.encoder_batch and self.encoder_batch > 1:
            for t in idxs:
                keep[t]["_static"], keep[t]["_gen"] = True, self._static_gen
        st["cached_features"] = keep
        return keep[frame_idx]

    def _encoder_graph(self, n, slot=0, partition=None):
        """Captured image-encoder pass over n frames -> (graph, static input [n,3,S,S], static outputs, #kernels).
        With `partition` the graph is captured on (and must be replayed from) that SM partition's stream, persistent
        kernels sized to its SM count; `slot` names one of several independent sets of static buffers."""
        key = ("encoder", n, slot, partition is not None)
        ent = self._graphs.get(key)
        if ent is not None:
            return ent
        eng = self.engine()
        static_in = torch.zeros((n, 3, self.image_size, self.image_size), dtype=torch.float32, device=self.device)
        stream = partition.stream if partition is not None else None
        try:
            # persistent grids: the partition's SMs, or the whole device for the full-device graph
            ops.set_sm_budget(partition.sms if partition is not None else 0)
            if partition is not None:
                stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(stream):
                    eng.encode_frames(static_in)
            else:
                eng.encode_frames(static_in)  # warm-up: one-time kernel attribute setup must not happen during capture
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            before = _lib.launch_count
            with torch.cuda.graph(graph, stream=stream):
                out = eng.encode_frames(static_in)
            n_kernels = _lib.launch_count - before
            _lib.launch_count = before
        finally:
            part = self._partition_obj  # tracked frames plan for the SMs the encoder partition leaves
            ops.set_sm_budget(part.total_sms - part.sms if part is not None else 0)
        ent = (graph, static_in, out, n_kernels)
        self._graphs[key] = ent
        return ent

    def _load_frames(self, st, idxs, static_in):
        """Copy the listed frames into a static encoder input (one copy when they are a contiguous ascending run)."""
        images = st["images"]
        if (torch.is_tensor(images) and images.is_cuda and images.dtype == torch.float32 and len(idxs) > 1
                and all(b - a == 1 for a, b in zip(idxs, idxs[1:]))):
            static_in[: len(idxs)].copy_(images[idxs[0]: idxs[-1] + 1], non_blocking=True)
            return
        for j, t in enumerate(idxs):
            static_in[j].copy_(self._frame(st, t), non_blocking=True)

    # ------------------------------------------------------------------ encoder running ahead of the propagation
    def _partition(self):
        """The encoder's SM partition (created on first use), or None when disabled / not available on this driver."""
        if self.encoder_sms <= 0 or not self.use_cuda_graphs or self._partition_error is not None:
            return None
        if self._partition_obj is None:
            from .pipeline import SmPartition

            try:
                part = SmPartition(self.device, self.encoder_sms)
            except Exception as e:  # no green contexts on this driver / cuda-python missing: alternate as before
                self._partition_error = repr(e)
                warnings.warn(f"encoder SM partition unavailable ({e!r}); encoder and tracking will alternate")
                return None
            self._partition_obj = part
            self.engine().sm_budget = part.total_sms - part.sms
            ops.set_sm_budget(part.total_sms - part.sms)
            # graphs captured so far assumed the whole device for the tracked frame
            self._graphs = {k: v for k, v in self._graphs.items() if k[0] == "encoder"}
            self._graph_seen = {}
        return self._partition_obj

    def attach_remote_encoders(self, remote):
        """Single clip on several GPUs: `remote` (pipeline.RemoteEncoders) names the ranks that run
        `serve_encoder`; propagate_in_video then receives every look-ahead batch from them instead of encoding it."""
        self._remote = remote

    @torch.inference_mode()
    def serve_encoder(self, images, my_index, num_encoders, dst=0, group=None, max_plans=None):
        """Encoder-rank side of the single-clip mode: encode the batches of each announced plan that fall to this rank
        and send their features to rank `dst`.  `images`: the normalised clip [T,3,S,S] resident on this GPU."""
        from .pipeline import serve_clip_encoder

        self._sync_engine()
        st = {"images": images, "num_frames": len(images)}
        eng = self.engine()

        def encode(frames, slot):
            if self.use_cuda_graphs and len(frames) == self.encoder_batch and self.encoder_batch > 1:
                graph, static_in, out, n_kernels = self._encoder_graph(len(frames), slot)
                self._load_frames(st, frames, static_in)
                graph.replay()
                _lib.launch_count += n_kernels
                return out
            imgs = torch.stack([self._frame(st, t) for t in frames]).contiguous()
            return eng.encode_frames(imgs)

        return serve_clip_encoder(encode, my_index, num_encoders, self.device, dst=dst, group=group,
                                  max_plans=max_plans)

    def _begin_pipeline(self, st, order, reverse):
        """Set up the look-ahead encoder for the frames `order` still has to track (None: nothing to overlap)."""
        from .pipeline import BatchPlan, FeaturePipeline, PartitionProducer, RemoteProducer

        done = self._frames_without_tracking(st)
        tracked = [t for t in order if t not in done]
        if not tracked:
            return None
        first, last, step, n = tracked[0], tracked[-1], (-1 if reverse else 1), self.encoder_batch
        if self._remote is not None:
            plan = BatchPlan(first, last, step, n, include_tail=True)
            self._remote.announce(plan)
            return FeaturePipeline(plan, RemoteProducer(self._remote, n, self.device, self.cfg.feat), depth=len(self._remote.ranks))
        if n < 2 or len(tracked) <= n:
            return None  # a single batch: nothing to overlap with
        if self._pipeline_owner is not None and self._pipeline_owner is not st:
            return None  # another session is mid-propagation and owns the look-ahead slots: alternate for this one
        part = self._partition()
        if part is None:
            return None
        # a final partial batch replays the same n-frame graph (stale inputs in the unused rows, outputs ignored)
        plan = BatchPlan(first, last, step, n, include_tail=True)
        try:
            self._encoder_graph(n)  # first batch of a pass: full device, nothing to overlap with
            for slot in range(2):
                self._encoder_graph(n, slot, part)
        except Exception as e:  # the partition exists but the encoder cannot run on it: alternate as before
            self._partition_error = repr(e)
            warnings.warn(f"image encoder could not be captured on the SM partition ({e!r}); alternating instead")
            return None
        self._pipeline_owner = st
        return FeaturePipeline(plan, PartitionProducer(self, st, part, n), depth=1)

    def _frame(self, st, t):
        img = st["images"][t]
        return img.to(self.device, non_blocking=True).float()

    # ------------------------------------------------------------------ prompts
    @torch.inference_mode()
    def add_new_points_or_box(self, inference_state, frame_idx, obj_id, points=None, labels=None,
                              clear_old_points=True, normalize_coords=True, box=None):
        """Add new points (and/or a box) to a frame (reference :173-314)."""
        st = inference_state
        self._sync_engine()
        obj_idx = self._obj_id_to_idx(st, obj_id)
        if (points is not None) != (labels is not None):
            raise ValueError("points and labels must be provided together")
        if points is None and box is None:
            raise ValueError("at least one of points or box must be provided as input")
        if points is None:
            points = torch.zeros(0, 2, dtype=torch.float32)
        elif not isinstance(points, torch.Tensor):
            points = torch.tensor(points, dtype=torch.float32)
        if labels is None:
            labels = torch.zeros(0, dtype=torch.int32)
        elif not isinstance(labels, torch.Tensor):
            labels = torch.tensor(labels, dtype=torch.int32)
        points = points.reshape(1, -1, 2).float().cpu()
        labels = labels.reshape(1, -1).to(torch.int32).cpu()
        if box is not None:
            if not clear_old_points:
                raise ValueError("cannot add box without clearing old points, since box prompt must be provided "
                                 "before any point prompt (please use clear_old_points=True instead)")
            if st["tracking_has_started"]:
                warnings.warn("You are adding a box after tracking starts. SAM 2 may not always be able to "
                              "incorporate a box prompt for *refinement*. If you intend to use box prompt as an "
                              "*initial* input before tracking, please call 'reset_state' on the inference state "
                              "to restart from scratch.", category=UserWarning, stacklevel=2)
            box = torch.as_tensor(box, dtype=torch.float32).cpu().reshape(1, 2, 2)
            points = torch.cat([box, points], dim=1)
            labels = torch.cat([torch.tensor([[2, 3]], dtype=torch.int32), labels], dim=1)
        if normalize_coords:
            points = points / torch.tensor([st["video_width"], st["video_height"]], dtype=torch.float32)
        points = (points * self.image_size).to(self.device)
        labels = labels.to(self.device)
        old = None if clear_old_points else st["point_inputs_per_obj"][obj_idx].get(frame_idx, None)
        if old is not None:
            points = torch.cat([old["point_coords"], points], dim=1)
            labels = torch.cat([old["point_labels"], labels], dim=1)
        point_inputs = {"point_coords": points, "point_labels": labels}
        st["point_inputs_per_obj"][obj_idx][frame_idx] = point_inputs
        st["mask_inputs_per_obj"][obj_idx].pop(frame_idx, None)

        tracked = self._tracked_info(st, obj_idx, frame_idx)
        is_init_cond_frame = tracked is None
        reverse = False if is_init_cond_frame else tracked["reverse"]
        obj_output_dict = st["output_dict_per_obj"][obj_idx]
        obj_temp_output_dict = st["temp_output_dict_per_obj"][obj_idx]
        is_cond = is_init_cond_frame or self.add_all_frames_to_correct_as_cond
        storage_key = "cond_frame_outputs" if is_cond else "non_cond_frame_outputs"
        prev_out = obj_temp_output_dict[storage_key].get(frame_idx)
        if prev_out is None:
            prev_out = obj_output_dict["cond_frame_outputs"].get(frame_idx)
            if prev_out is None:
                prev_out = obj_output_dict["non_cond_frame_outputs"].get(frame_idx)
        prev_sam_mask_logits = None
        if prev_out is not None and prev_out["pred_masks"] is not None:
            prev_sam_mask_logits = torch.clamp(prev_out["pred_masks"].to(self.device), -32.0, 32.0)
        current_out, _ = self._run_single_frame_inference(
            st, obj_output_dict, frame_idx, 1, is_init_cond_frame, point_inputs, None, reverse,
            run_mem_encoder=False, prev_sam_mask_logits=prev_sam_mask_logits, obj0=obj_idx)
        obj_temp_output_dict[storage_key][frame_idx] = current_out
        consolidated = self._consolidate_temp_output_across_obj(st, frame_idx, is_cond, run_mem_encoder=False,
                                                                consolidate_at_video_res=True)
        _, video_res_masks = self._get_orig_video_res_output(st, consolidated["pred_masks_video_res"])
        return frame_idx, st["obj_ids"], video_res_masks

    def add_new_points(self, *args, **kwargs):
        """Deprecated alias (reference :316-318)."""
        return self.add_new_points_or_box(*args, **kwargs)

    @torch.inference_mode()
    def add_new_mask(self, inference_state, frame_idx, obj_id, mask):
        """Add new mask to a frame (reference :321-402)."""
        st = inference_state
        self._sync_engine()
        obj_idx = self._obj_id_to_idx(st, obj_id)
        if not isinstance(mask, torch.Tensor):
            mask = torch.tensor(mask, dtype=torch.bool)
        assert mask.dim() == 2
        mask_H, mask_W = mask.shape
        m = mask[None, None].float().to(self.device)
        if mask_H != self.image_size or mask_W != self.image_size:
            m = ops.resize_bilinear_aa(m, self.image_size, self.image_size, binarize_half=True)
        st["mask_inputs_per_obj"][obj_idx][frame_idx] = m
        st["point_inputs_per_obj"][obj_idx].pop(frame_idx, None)
        tracked = self._tracked_info(st, obj_idx, frame_idx)
        is_init_cond_frame = tracked is None
        reverse = False if is_init_cond_frame else tracked["reverse"]
        obj_output_dict = st["output_dict_per_obj"][obj_idx]
        obj_temp_output_dict = st["temp_output_dict_per_obj"][obj_idx]
        is_cond = is_init_cond_frame or self.add_all_frames_to_correct_as_cond
        storage_key = "cond_frame_outputs" if is_cond else "non_cond_frame_outputs"
        current_out, _ = self._run_single_frame_inference(
            st, obj_output_dict, frame_idx, 1, is_init_cond_frame, None, m, reverse, run_mem_encoder=False,
            obj0=obj_idx)
        obj_temp_output_dict[storage_key][frame_idx] = current_out
        consolidated = self._consolidate_temp_output_across_obj(st, frame_idx, is_cond, run_mem_encoder=False,
                                                                consolidate_at_video_res=True)
        _, video_res_masks = self._get_orig_video_res_output(st, consolidated["pred_masks_video_res"])
        return frame_idx, st["obj_ids"], video_res_masks

    # ------------------------------------------------------------------ outputs
    def _get_orig_video_res_output(self, st, any_res_masks):
        """Resize to the original video resolution (reference :404-424)."""
        vh, vw = st["video_height"], st["video_width"]
        any_res_masks = any_res_masks.to(self.device)
        if tuple(any_res_masks.shape[-2:]) == (vh, vw):
            video_res_masks = any_res_masks
        else:
            video_res_masks = ops.resize_bilinear(any_res_masks, vh, vw)
        if self.non_overlap_masks:
            video_res_masks = self._apply_non_overlapping_constraints(video_res_masks)
        return any_res_masks, video_res_masks

    @staticmethod
    def _apply_non_overlapping_constraints(pred_masks):
        """Keep only the highest-scoring object per pixel (sam2_base.py:1663-1681).  Optional post-step, off in
        the shipped config (usvm_non_overlap_f32)."""
        if pred_masks.size(0) == 1:
            return pred_masks
        return ops.non_overlap(pred_masks)

    def _mem_view(self, mem_tok):
        """token-major bf16 [B,T,64] -> the reference's [B,64,feat,feat] layout (a view)."""
        B, fs = mem_tok.shape[0], self.cfg.feat
        return mem_tok.view(B, fs, fs, self.mem_dim).permute(0, 3, 1, 2)

    def _maskmem_pos_enc(self, st, B):
        """Cached constant, expanded per object (reference :1016-1039)."""
        c = st["constants"]
        if "maskmem_pos_enc" not in c:
            fs = self.cfg.feat
            pos = self.engine().w.mem_pos.view(1, fs, fs, self.mem_dim).permute(0, 3, 1, 2)
            c["maskmem_pos_enc"] = [pos]
        return [x.expand(B, -1, -1, -1) for x in c["maskmem_pos_enc"]]

    def _consolidate_temp_output_across_obj(self, st, frame_idx, is_cond, run_mem_encoder,
                                            consolidate_at_video_res=False):
        """Merge per-object temporary outputs of a frame into one batched output (reference :426-554)."""
        B = self._get_obj_num(st)
        storage_key = "cond_frame_outputs" if is_cond else "non_cond_frame_outputs"
        if consolidate_at_video_res:
            assert not run_mem_encoder, "memory encoder cannot run at video resolution"
            ch, cw, mask_key = st["video_height"], st["video_width"], "pred_masks_video_res"
        else:
            ch = cw = self.image_size // 4
            mask_key = "pred_masks"
        dev = self.device
        out = {"maskmem_features": None, "maskmem_pos_enc": None, "_mem_tok": None,
               mask_key: torch.full((B, 1, ch, cw), NO_OBJ_SCORE, dtype=torch.float32, device=dev),
               "obj_ptr": torch.full((B, self.hidden_dim), NO_OBJ_SCORE, dtype=torch.float32, device=dev),
               "object_score_logits": torch.full((B, 1), 10.0, dtype=torch.float32, device=dev)}
        empty_mask_ptr = None
        for obj_idx in range(B):
            tmp = st["temp_output_dict_per_obj"][obj_idx]
            perm = st["output_dict_per_obj"][obj_idx]
            o = tmp[storage_key].get(frame_idx, None)
            if o is None:
                o = perm["cond_frame_outputs"].get(frame_idx, None)
            if o is None:
                o = perm["non_cond_frame_outputs"].get(frame_idx, None)
            if o is None:
                if run_mem_encoder:
                    if empty_mask_ptr is None:
                        empty_mask_ptr = self._get_empty_mask_ptr(st, frame_idx)
                    out["obj_ptr"][obj_idx:obj_idx + 1] = empty_mask_ptr
                continue
            m = o["pred_masks"]
            if tuple(m.shape[-2:]) != (ch, cw):
                m = ops.resize_bilinear(m, ch, cw)
            out[mask_key][obj_idx:obj_idx + 1] = m
            out["obj_ptr"][obj_idx:obj_idx + 1] = o["obj_ptr"]
            out["object_score_logits"][obj_idx:obj_idx + 1] = o["object_score_logits"]
        if run_mem_encoder:
            mem_tok = self._run_memory_encoder(st, frame_idx, B, out["pred_masks"], out["object_score_logits"], True)
            out["_mem_tok"] = mem_tok
            out["maskmem_features"] = self._mem_view(mem_tok)
            out["maskmem_pos_enc"] = self._maskmem_pos_enc(st, B)
        return out

    def _get_empty_mask_ptr(self, st, frame_idx):
        """Dummy object pointer from an empty mask (reference :556-590)."""
        eng = self.engine()
        f = self._get_image_feature(st, frame_idx)
        z = torch.zeros((1, 1, self.image_size, self.image_size), dtype=torch.float32, device=self.device)
        return eng.mask_as_output(f["feat"], f["feat_s0"], f["feat_s1"], z, 1)["obj_ptr"]

    # ------------------------------------------------------------------ propagation
    @torch.inference_mode()
    def propagate_in_video_preflight(self, inference_state):
        """Consolidate temporary outputs before tracking (reference :593-660)."""
        st = inference_state
        self._sync_engine()
        st["tracking_has_started"] = True
        B = self._get_obj_num(st)
        self._ensure_store(st)
        temp = st["temp_output_dict_per_obj"]
        output_dict = st["output_dict"]
        cfi = st["consolidated_frame_inds"]
        for is_cond in (False, True):
            storage_key = "cond_frame_outputs" if is_cond else "non_cond_frame_outputs"
            temp_frame_inds = set()
            for obj_temp in temp.values():
                temp_frame_inds.update(obj_temp[storage_key].keys())
            cfi[storage_key].update(temp_frame_inds)
            for frame_idx in temp_frame_inds:
                consolidated = self._consolidate_temp_output_across_obj(st, frame_idx, is_cond, run_mem_encoder=True)
                consolidated = self._commit(st, frame_idx, consolidated)
                output_dict[storage_key][frame_idx] = consolidated
                self._add_output_per_object(st, frame_idx, consolidated, storage_key)
                if self.clear_non_cond_mem_around_input and (self.clear_non_cond_mem_for_multi_obj or B <= 1):
                    self._clear_non_cond_mem_around_input(st, frame_idx)
            for obj_temp in temp.values():
                obj_temp[storage_key].clear()
        for frame_idx in output_dict["cond_frame_outputs"]:
            output_dict["non_cond_frame_outputs"].pop(frame_idx, None)
        for obj_out in st["output_dict_per_obj"].values():
            for frame_idx in obj_out["cond_frame_outputs"]:
                obj_out["non_cond_frame_outputs"].pop(frame_idx, None)
        for frame_idx in cfi["cond_frame_outputs"]:
            assert frame_idx in output_dict["cond_frame_outputs"]
            cfi["non_cond_frame_outputs"].discard(frame_idx)
        # the device control block of a tracked frame names every selected memory frame and object pointer
        # (include/usvm2_b200.h: usvm_frame_ctrl); the reference has no such bound (sam2_base.py:1296-1394)
        n_cond = len(output_dict["cond_frame_outputs"])
        if self.cfg.max_cond_frames_in_attn != -1:
            n_cond = min(n_cond, self.cfg.max_cond_frames_in_attn)
        if (n_cond + self.num_maskmem - 1 > _lib.MAX_MEMORY_FRAMES
                or n_cond + self.cfg.max_obj_ptrs_in_encoder - 1 > _lib.MAX_PTRS):
            lim = min(_lib.MAX_MEMORY_FRAMES - (self.num_maskmem - 1), _lib.MAX_PTRS - (self.cfg.max_obj_ptrs_in_encoder - 1))
            raise RuntimeError(f"{n_cond} conditioning (prompted) frames take part in the memory attention, this build's "
                               f"frame control block holds at most {lim}; set max_cond_frames_in_attn <= {lim} "
                               "(++model.max_cond_frames_in_attn) to keep the closest ones like the reference does")
        all_consolidated = cfi["cond_frame_outputs"] | cfi["non_cond_frame_outputs"]
        input_frames = set()
        for d in st["point_inputs_per_obj"].values():
            input_frames.update(d.keys())
        for d in st["mask_inputs_per_obj"].values():
            input_frames.update(d.keys())
        assert all_consolidated == input_frames

    @torch.inference_mode()
    def propagate_in_video(self, inference_state, start_frame_idx=None, max_frame_num_to_track=None, reverse=False):
        """Propagate the prompts across the video; generator of (frame_idx, obj_ids, video_res_masks)
        (reference :663-745)."""
        st = inference_state
        self.propagate_in_video_preflight(st)
        output_dict = st["output_dict"]
        cfi = st["consolidated_frame_inds"]
        obj_ids = st["obj_ids"]
        num_frames = st["num_frames"]
        B = self._get_obj_num(st)
        cond_frames = self._cond_frames(st)
        if len(cond_frames) == 0:
            raise RuntimeError("No points are provided; please add points first")
        clear_non_cond_mem = self.clear_non_cond_mem_around_input and (self.clear_non_cond_mem_for_multi_obj or B <= 1)
        if start_frame_idx is None:
            start_frame_idx = min(cond_frames)
        if max_frame_num_to_track is None:
            max_frame_num_to_track = num_frames
        if reverse:
            end_frame_idx = max(start_frame_idx - max_frame_num_to_track, 0)
            order = range(start_frame_idx, end_frame_idx - 1, -1) if start_frame_idx > 0 else []
        else:
            end_frame_idx = min(start_frame_idx + max_frame_num_to_track, num_frames - 1)
            order = range(start_frame_idx, end_frame_idx + 1)
        st["_pipeline"] = self._begin_pipeline(st, list(order), reverse)
        st["_pass_last"] = end_frame_idx  # (frame pipelining stops at the last frame of the pass)
        self._pipe_state = None
        try:
            yield from self._propagate_loop(st, order, reverse, clear_non_cond_mem, B)
        finally:
            st.pop("_pass_last", None)
            self._pipe_state = None
            pipe = st.pop("_pipeline", None)
            if pipe is not None:
                pipe.close()
            if self._pipeline_owner is st:
                self._pipeline_owner = None

    # ------------------------------------------------------------------ several sessions in lock-step
    @torch.inference_mode()
    def propagate_in_videos(self, inference_states, start_frame_idx=None, max_frame_num_to_track=None, reverse=False):
        """Propagate S independent sessions in LOCK-STEP as one batched frame: generator of
        (frame_idx, [obj_ids of each session], [video_res_masks of each session]).

        The reference's only inference parallelism is a process pool over volumes
        (medsam2_infer_CT_lesion_npz_recist.py:454-456); on one B200 a single object's tracked frame is a latency chain
        that leaves most of the device idle, so independent videos are stacked into the batch dimension of ONE captured
        frame graph instead: S sessions x Bo objects run as B = S * Bo objects whose per-video operands (backbone
        features, pix_feat projection) are indexed by object // Bo.  The sessions share one frame store (each session's
        store becomes a column view of it) and one device control block -- they advance through the same frame indices
        with the same memory-bank layout, which requires: equal frame count, video size and object count, prompts on the
        same frames, and no differing earlier tracking results.  Anything else raises ValueError (track such sessions
        one after another with propagate_in_video).  No host synchronisation happens per frame; every session's
        `inference_state` ends up exactly as if it had been tracked alone (results are views of the shared store)."""
        states = list(inference_states)
        if len(states) == 0:
            return
        for st in states:
            self.propagate_in_video_preflight(st)
        self._check_lockstep(states)
        st0 = states[0]
        S, Bo = len(states), self._get_obj_num(st0)
        num_frames = st0["num_frames"]
        cond_frames = self._cond_frames(st0)
        if len(cond_frames) == 0:
            raise RuntimeError("No points are provided; please add points first")
        if start_frame_idx is None:
            start_frame_idx = min(cond_frames)
        if max_frame_num_to_track is None:
            max_frame_num_to_track = num_frames
        if reverse:
            end_frame_idx = max(start_frame_idx - max_frame_num_to_track, 0)
            order = range(start_frame_idx, end_frame_idx - 1, -1) if start_frame_idx > 0 else []
        else:
            end_frame_idx = min(start_frame_idx + max_frame_num_to_track, num_frames - 1)
            order = range(start_frame_idx, end_frame_idx + 1)
        shared = self._merge_stores(states)
        feats = _LockstepFeatures(self, states, -1 if reverse else 1)
        hw = (st0["video_height"], st0["video_width"])
        clear_non_cond_mem = self.clear_non_cond_mem_around_input and (self.clear_non_cond_mem_for_multi_obj or Bo <= 1)
        cfi = st0["consolidated_frame_inds"]
        for frame_idx in _progress(order, "propagate in videos"):
            if frame_idx in cfi["cond_frame_outputs"] or frame_idx in cfi["non_cond_frame_outputs"]:
                storage_key = "cond_frame_outputs" if frame_idx in cfi["cond_frame_outputs"] else "non_cond_frame_outputs"
                masks = []
                for st in states:
                    current_out = st["output_dict"][storage_key][frame_idx]
                    if clear_non_cond_mem and storage_key == "cond_frame_outputs":
                        self._clear_non_cond_mem_around_input(st, frame_idx)
                    self._add_output_per_object(st, frame_idx, current_out, storage_key)
                    masks.append(self._get_orig_video_res_output(st, current_out["pred_masks"])[1])
            else:
                storage_key = "non_cond_frame_outputs"
                f = feats.get(frame_idx)
                mem_inputs = self._memory_inputs(st0, frame_idx, st0["output_dict"], reverse)
                video = self._run_tracked_frame(shared, f, frame_idx, S * Bo, mem_inputs, hw, group=Bo)
                masks = []
                for i, st in enumerate(states):
                    current_out = self._slot_views(st, frame_idx)
                    st["output_dict"][storage_key][frame_idx] = current_out
                    self._add_output_per_object(st, frame_idx, current_out, storage_key)
                    m = video[i * Bo:(i + 1) * Bo]
                    masks.append(self._apply_non_overlapping_constraints(m) if self.non_overlap_masks else m)
            for st in states:
                st["frames_already_tracked"][frame_idx] = {"reverse": reverse}
                for tracked in st.get("frames_tracked_per_obj", {}).values():
                    tracked[frame_idx] = {"reverse": reverse}
            yield frame_idx, [st["obj_ids"] for st in states], masks

    def _check_lockstep(self, states):
        st0 = states[0]

        def signature(st):
            od = st["output_dict"]
            return (st["num_frames"], st["video_height"], st["video_width"], self._get_obj_num(st),
                    tuple(sorted(od["cond_frame_outputs"])), tuple(sorted(od["non_cond_frame_outputs"])),
                    bool(st.get("_per_object")))

        ref = signature(st0)
        for i, st in enumerate(states[1:], 1):
            if signature(st) != ref:
                raise ValueError(f"session {i} cannot be tracked in lock-step with session 0: frame count, video size, "
                                 "object count, prompted frames and earlier tracked frames must agree "
                                 f"({signature(st)} vs {ref}); use propagate_in_video for it")
        if ref[3] == 0:
            raise RuntimeError("No points are provided; please add points first")
        if ref[6]:
            raise ValueError("sessions on the per-object (EfficientTAM) schedule cannot be batched")
        if len({id(st) for st in states}) != len(states):
            raise ValueError("the same inference_state was passed twice")

    def _merge_stores(self, states):
        """One frame store for all sessions (objects of session i in columns [i * Bo, (i + 1) * Bo)); each session's own
        store becomes a column view of it, its stored entries are re-pointed."""
        st0 = states[0]
        S, Bo, T = len(states), self._get_obj_num(st0), st0["num_frames"]
        owners = [st["_store"] for st in states]
        base = getattr(owners[0], "shared", None)
        if base is not None and base.B == S * Bo and all(getattr(o, "shared", None) is base and o.column0 == i * Bo
                                                         for i, o in enumerate(owners)):
            return base  # already merged by an earlier pass over the same sessions
        shared = ops.FrameStore(T, S * Bo, self.device, T=self.cfg.feat ** 2, low=self.image_size // 4)
        for i, st in enumerate(states):
            old = st["_store"]
            od = st["output_dict"]
            for t in list(od["cond_frame_outputs"]) + list(od["non_cond_frame_outputs"]):
                lo = i * Bo
                shared.mem[t, lo:lo + Bo].copy_(old.mem[t])
                shared.ptr[t, lo:lo + Bo].copy_(old.ptr[t])
                shared.score[t, lo:lo + Bo].copy_(old.score[t])
                shared.masks[t, lo:lo + Bo].copy_(old.masks[t])
            st["_store"] = shared.columns(i * Bo, Bo)
            self._refresh_views(st)
        return shared

    def _cond_frames(self, st):
        """Frames that hold a conditioning output (the default start of a pass is the earliest one, reference :684-686)."""
        return set(st["output_dict"]["cond_frame_outputs"])

    def _frames_without_tracking(self, st):
        """Frames a pass will not run the tracking step on (their consolidated outputs exist already)."""
        cfi = st["consolidated_frame_inds"]
        return cfi["cond_frame_outputs"] | cfi["non_cond_frame_outputs"]

    def _propagate_loop(self, st, order, reverse, clear_non_cond_mem, B):
        output_dict = st["output_dict"]
        cfi = st["consolidated_frame_inds"]
        obj_ids = st["obj_ids"]
        for frame_idx in _progress(order, "propagate in video"):
            if frame_idx in cfi["cond_frame_outputs"]:
                storage_key = "cond_frame_outputs"
                current_out = output_dict[storage_key][frame_idx]
                pred_masks = current_out["pred_masks"]
                if clear_non_cond_mem:
                    self._clear_non_cond_mem_around_input(st, frame_idx)
            elif frame_idx in cfi["non_cond_frame_outputs"]:
                storage_key = "non_cond_frame_outputs"
                current_out = output_dict[storage_key][frame_idx]
                pred_masks = current_out["pred_masks"]
            else:
                storage_key = "non_cond_frame_outputs"
                current_out, video_res_masks = self._track_frame(st, frame_idx, B, reverse, output_dict)
                output_dict[storage_key][frame_idx] = current_out
                pred_masks = None
            self._add_output_per_object(st, frame_idx, current_out, storage_key)
            st["frames_already_tracked"][frame_idx] = {"reverse": reverse}
            for tracked in st.get("frames_tracked_per_obj", {}).values():  # EfficientTAM's per-object bookkeeping
                tracked[frame_idx] = {"reverse": reverse}
            if pred_masks is not None:
                _, video_res_masks = self._get_orig_video_res_output(st, pred_masks)
            elif self.non_overlap_masks:
                video_res_masks = self._apply_non_overlapping_constraints(video_res_masks)
            yield frame_idx, obj_ids, video_res_masks

    def _add_output_per_object(self, st, frame_idx, current_out, storage_key):
        """Per-object views of a batched output (reference :747-774)."""
        mem = current_out["maskmem_features"]
        pos = current_out["maskmem_pos_enc"]
        for obj_idx, obj_output_dict in st["output_dict_per_obj"].items():
            s = slice(obj_idx, obj_idx + 1)
            o = {"maskmem_features": None, "maskmem_pos_enc": None, "_mem_tok": None,
                 "_slot": current_out.get("_slot"),
                 "pred_masks": current_out["pred_masks"][s], "obj_ptr": current_out["obj_ptr"][s],
                 "object_score_logits": current_out["object_score_logits"][s]}
            if mem is not None:
                o["maskmem_features"] = mem[s]
                o["_mem_tok"] = current_out["_mem_tok"][s]
            if pos is not None:
                o["maskmem_pos_enc"] = [x[s] for x in pos]
            obj_output_dict[storage_key][frame_idx] = o

    # ------------------------------------------------------------------ one frame
    def _memory_inputs(self, st, frame_idx, output_dict, reverse):
        """Memory-bank selection (sam2_base.py:1296-1394) as store slots:
        (mem_slots, tpos_rows, ptr_slots, ptr_rel)."""
        cfg = self.cfg
        cond = output_dict["cond_frame_outputs"]
        assert len(cond) > 0
        selected, unselected = _select_closest_cond_frames(frame_idx, cond, cfg.max_cond_frames_in_attn)
        entries = [(0, o) for o in selected.values()]
        r = cfg.memory_temporal_stride_for_eval
        non_cond = output_dict["non_cond_frame_outputs"]
        for t_pos in range(1, cfg.num_maskmem):
            t_rel = cfg.num_maskmem - t_pos
            if t_rel == 1:
                prev = frame_idx + t_rel if reverse else frame_idx - t_rel
            elif not reverse:
                prev = ((frame_idx - 2) // r) * r - (t_rel - 2) * r
            else:
                prev = -(-(frame_idx + 2) // r) * r + (t_rel - 2) * r
            o = non_cond.get(prev, None)
            if o is None:
                o = unselected.get(prev, None)
            entries.append((t_pos, o))
        mem_slots, tpos_rows = [], []
        for t_pos, o in entries:
            if o is None:
                continue
            mem_slots.append(o["_slot"])
            tpos_rows.append(cfg.num_maskmem - t_pos - 1)
        num_frames = st["num_frames"]
        max_ptrs = min(num_frames, cfg.max_obj_ptrs_in_encoder)
        sign = -1 if reverse else 1
        denom = float(max_ptrs - 1)
        ptr_slots, ptr_rel = [], []
        for t, o in selected.items():
            if (t >= frame_idx) if reverse else (t <= frame_idx):
                ptr_slots.append(o["_slot"])
                ptr_rel.append((frame_idx - t) * sign / denom)
        for t_diff in range(1, max_ptrs):
            t = frame_idx + t_diff if reverse else frame_idx - t_diff
            if t < 0 or t >= num_frames:
                break
            o = non_cond.get(t, unselected.get(t, None))
            if o is not None:
                ptr_slots.append(o["_slot"])
                ptr_rel.append(t_diff / denom)
        return mem_slots, tpos_rows, ptr_slots, ptr_rel

    def _track_frame(self, st, frame_idx, B, reverse, output_dict, obj0=None):
        """One tracked (unprompted) frame of all B objects: results go straight into the frame store; the steady
        state is replayed from a CUDA graph (one graph per (B, #memories, #pointers, video size) signature).
        obj0: track ONE object (B == 1) of a multi-object session against its own `output_dict` -- the control block
        addresses that object's column of the store."""
        look = -1 if reverse else 1
        f = self._get_image_feature(st, frame_idx, lookahead=look)
        mem_inputs = self._memory_inputs(st, frame_idx, output_dict, reverse)
        hw = (st["video_height"], st["video_width"])
        nxt = None
        if obj0 is None and self.pipeline_frames and self.use_cuda_graphs:
            # the frame tracked next (if it is the neighbour and still untracked): its features feed the pipelined prefix
            t1 = frame_idx + look
            cfi = st["consolidated_frame_inds"]
            if (0 <= t1 < st["num_frames"] and t1 not in cfi["cond_frame_outputs"]
                    and t1 not in cfi["non_cond_frame_outputs"] and frame_idx != st.get("_pass_last", frame_idx)):
                nxt = (t1, lambda: self._get_image_feature(st, t1, lookahead=look))
        video = self._run_tracked_frame(st["_store"], f, frame_idx, B, mem_inputs, hw, obj0=obj0 or 0, nxt=nxt)
        if obj0 is not None:
            return self._obj_slot_views(st, frame_idx, obj0), video
        return self._slot_views(st, frame_idx), video

    def _run_tracked_frame(self, store, f, frame_idx, B, mem_inputs, hw, obj0=0, group=0, nxt=None):
        """Enqueue one tracked frame of B objects whose results land in slot `frame_idx` of `store` (columns obj0 ...
        obj0 + B - 1); returns the video-resolution logits [B,1,H,W].  group > 0: B / group videos of `group` objects
        each in lock-step (propagate_in_videos), f holds one frame of features per video.
        nxt = (index of the frame tracked next, callable returning its features): steady-state frames then run the
        pipelined graphs (see _run_pipelined_frame)."""
        eng = self.engine()
        mem_slots, tpos_rows, ptr_slots, ptr_rel = mem_inputs
        key = (B, len(mem_slots), len(ptr_slots), hw, self.fill_hole_area, group)
        steady = self._last_frame_key == key  # (ramp-up signatures change every frame: not worth two more graphs each)
        self._last_frame_key = key
        if nxt is not None and group == 0 and steady and key in self._graphs:
            return self._run_pipelined_frame(store, f, frame_idx, key, mem_inputs, obj0, nxt)
        self._pipe_state = None
        ent = self._graphs.get(key) if self.use_cuda_graphs else None
        if self.use_cuda_graphs and ent is None:
            seen = self._graph_seen.get(key, 0) + 1
            self._graph_seen[key] = seen
            # a signature that recurs is worth a graph: the steady state (full bank) is captured at its second frame, the
            # ramp-up signatures of a clip (bank filling up) when a second clip / pass reaches them
            if seen >= 2:
                ops.set_frame_ctrl(self._ctrl, store, obj0, frame_idx, mem_slots, tpos_rows, ptr_slots, ptr_rel)
                ent = self._capture_graph(key, f)
        if ent is not None:
            graph, static_f, video, n_kernels = ent
            # one launch refreshes the control block and copies this frame's features into the graph's static inputs
            ops.set_frame_ctrl(self._ctrl, store, obj0, frame_idx, mem_slots, tpos_rows, ptr_slots, ptr_rel,
                               copies=[(f[k], v) for k, v in static_f.items()])
            graph.replay()
            _lib.launch_count += n_kernels  # kernels of this library replayed by the graph
            return video.clone()
        ops.set_frame_ctrl(self._ctrl, store, obj0, frame_idx, mem_slots, tpos_rows, ptr_slots, ptr_rel)
        video, _ = eng.track_frame(f, self._ctrl, B, len(mem_slots), len(ptr_slots), hw, self.fill_hole_area, group=group)
        return video

    def _run_pipelined_frame(self, store, f, frame_idx, key, mem_inputs, obj0, nxt):
        """Steady-state frame t with the next frame's features at hand.  Two feature sets F[0], F[1] and two prefix sets
        P[0], P[1] (static buffers per graph signature) alternate by frame parity p: the graph of frame t reads F[p]
        (+ P[p], this frame's attention prefix, if the previous replay produced it: mode "pipe"; otherwise it computes the
        prefix inline: mode "start") and, on a forked branch, writes frame t + 1's prefix computed from F[p ^ 1] into
        P[p ^ 1].  A frame's features are copied into their set once -- one step earlier than without pipelining."""
        eng = self.engine()
        mem_slots, tpos_rows, ptr_slots, ptr_rel = mem_inputs
        B, n_mem, n_ptr, hw, fill, group = key
        bufs = self._pipe_bufs.get(key)
        if bufs is None:
            T = self.cfg.feat ** 2
            dev = self.device
            bufs = {"F": [{k: torch.empty_like(f[k]) for k in ("feat", "feat_bf16", "feat_s0", "feat_s1")} for _ in range(2)],
                    "P": [(torch.empty((B * T, 256), dtype=torch.float32, device=dev),
                           torch.empty((B * T, 256), dtype=torch.bfloat16, device=dev)) for _ in range(2)]}
            self._pipe_bufs[key] = bufs
        ps = self._pipe_state
        have = (ps is not None and ps["key"] == key and ps["frame"] == frame_idx and ps["store"] is store
                and ps["obj0"] == obj0)
        p = ps["parity"] if have else 0
        mode = "pipe" if have else "start"
        gkey = key + (mode, p)
        ent = self._graphs.get(gkey)
        t1, next_features = nxt
        F, P = bufs["F"], bufs["P"]
        if not have:  # this frame's features were not staged by a previous step
            ops.set_frame_ctrl(self._ctrl, store, obj0, frame_idx, mem_slots, tpos_rows, ptr_slots, ptr_rel,
                               copies=[(f[k], v) for k, v in F[p].items()])
        f1 = next_features()  # (only now: asking for it may hand this frame's encoder slot back to the producer)
        ops.set_frame_ctrl(self._ctrl, store, obj0, frame_idx, mem_slots, tpos_rows, ptr_slots, ptr_rel,
                           copies=[(f1[k], v) for k, v in F[p ^ 1].items()])
        if ent is None:
            eng.token_constants(B)
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            before = _lib.launch_count
            with torch.cuda.graph(graph):
                video, _ = eng.track_frame(F[p], self._ctrl, B, n_mem, n_ptr, hw, fill, group=group,
                                           prefix=P[p] if have else None, next_feat=F[p ^ 1]["feat"],
                                           next_prefix_out=P[p ^ 1])
            n_kernels = _lib.launch_count - before
            _lib.launch_count = before
            ent = (graph, None, video, n_kernels)
            self._graphs[gkey] = ent
        graph, _, video, n_kernels = ent
        graph.replay()
        _lib.launch_count += n_kernels
        self._pipe_state = dict(key=key, frame=t1, store=store, obj0=obj0, parity=p ^ 1)
        return video.clone()

    def _capture_graph(self, key, f, stream=None):
        B, n_mem, n_ptr, hw, fill, group = key
        eng = self.engine()
        static_f = {k: f[k].clone() for k in ("feat", "feat_bf16", "feat_s0", "feat_s1")}
        eng.token_constants(B)  # one-off constants must exist before the capture starts
        graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        before = _lib.launch_count
        with torch.cuda.graph(graph, stream=stream):
            video, _ = eng.track_frame(static_f, self._ctrl, B, n_mem, n_ptr, hw, fill, group=group)
        n_kernels = _lib.launch_count - before
        _lib.launch_count = before  # capturing launches nothing
        ent = (graph, static_f, video, n_kernels)
        self._graphs[key] = ent
        return ent

    def _run_single_frame_inference(self, st, output_dict, frame_idx, batch_size, is_init_cond_frame, point_inputs,
                                    mask_inputs, reverse, run_mem_encoder, prev_sam_mask_logits=None, obj0=0):
        """track_step + compact output for PROMPTED frames and per-object interaction
        (reference :912-978, sam2_base.py:1500-1651).  Unprompted tracking of all objects goes through _track_frame."""
        eng = self.engine()
        cfg = self.cfg
        B = batch_size
        assert point_inputs is None or mask_inputs is None
        f = self._get_image_feature(st, frame_idx)
        if mask_inputs is not None:
            o = eng.mask_as_output(f["feat"], f["feat_s0"], f["feat_s1"], mask_inputs, B)
        else:
            if is_init_cond_frame:
                T = cfg.feat ** 2
                pix, _ = ops.axpby(f["feat"], eng.w.no_mem_embed, rows=B * T, x_mod=T, y_mod=1)
            else:
                # correction clicks on an already tracked frame: condition on this object's memories in the store
                mem_slots, tpos_rows, ptr_slots, ptr_rel = self._memory_inputs(st, frame_idx, output_dict, reverse)
                ops.set_frame_ctrl(self._ctrl, st["_store"], obj0, frame_idx, mem_slots, tpos_rows, ptr_slots, ptr_rel)
                k_in, v_in, Nk, n_tok = eng.assemble_memory(self._ctrl, B, len(mem_slots), len(ptr_slots))
                pix = eng.memory_attention(f["feat"], k_in, v_in, Nk, n_tok, B)
            n_pts = 0 if point_inputs is None else point_inputs["point_labels"].size(1)
            multimask = (cfg.multimask_output_in_sam
                         and (is_init_cond_frame or cfg.multimask_output_for_tracking)
                         and (cfg.multimask_min_pt_num <= n_pts <= cfg.multimask_max_pt_num))
            if point_inputs is not None:
                sparse = eng.embed_points(point_inputs["point_coords"], point_inputs["point_labels"])
            else:
                sparse = eng.no_point_tokens(B)
            dense = None
            if prev_sam_mask_logits is not None:
                assert point_inputs is not None
                m = prev_sam_mask_logits
                L = self.image_size // 4
                if tuple(m.shape[-2:]) != (L, L):
                    m = ops.resize_bilinear_aa(m.float(), L, L)
                dense = eng.embed_mask_prompt(m.contiguous(), B)
            o = eng.sam_heads(pix, f["feat_s0"], f["feat_s1"], B, sparse, dense=dense, multimask=multimask)
        low = o["low"]
        mem_tok = None
        if run_mem_encoder and cfg.num_maskmem > 0:
            binarize = cfg.binarize_mask_from_pts_for_mem_enc and (point_inputs is not None)
            src = o.get("high", None)
            if src is None:
                src = low
            mask_in = eng.mem_mask_input(src, binarize, non_overlap=self.non_overlap_masks_for_mem_enc)
            mem_tok = eng.encode_memory(f["feat_bf16"], mask_in, o["score"], B)
        pred_masks = low
        if self.fill_hole_area > 0:
            pred_masks = ops.fill_holes(low, self.fill_hole_area)
        compact = {"maskmem_features": None if mem_tok is None else self._mem_view(mem_tok),
                   "maskmem_pos_enc": None if mem_tok is None else self._maskmem_pos_enc(st, B),
                   "_mem_tok": mem_tok, "pred_masks": pred_masks, "obj_ptr": o["obj_ptr"],
                   "object_score_logits": o["score"]}
        return compact, pred_masks

    def _run_memory_encoder(self, st, frame_idx, batch_size, masks, object_score_logits, is_mask_from_pts):
        """Memory encoder on consolidated masks (reference :980-1014); `masks` are low-res or 512^2 logits."""
        eng = self.engine()
        f = self._get_image_feature(st, frame_idx)
        binarize = self.cfg.binarize_mask_from_pts_for_mem_enc and is_mask_from_pts
        mask_in = eng.mem_mask_input(masks, binarize, non_overlap=self.non_overlap_masks_for_mem_enc)
        return eng.encode_memory(f["feat_bf16"], mask_in, object_score_logits, batch_size)

    # ------------------------------------------------------------------ editing surface
    @torch.inference_mode()
    def clear_all_prompts_in_frame(self, inference_state, frame_idx, obj_id, need_output=True):
        """Remove all prompts of an object on a frame (reference :776-846)."""
        st = inference_state
        obj_idx = self._obj_id_to_idx(st, obj_id)
        st["point_inputs_per_obj"][obj_idx].pop(frame_idx, None)
        st["mask_inputs_per_obj"][obj_idx].pop(frame_idx, None)
        temp = st["temp_output_dict_per_obj"]
        temp[obj_idx]["cond_frame_outputs"].pop(frame_idx, None)
        temp[obj_idx]["non_cond_frame_outputs"].pop(frame_idx, None)
        B = self._get_obj_num(st)
        frame_has_input = any(frame_idx in st["point_inputs_per_obj"][i] or frame_idx in st["mask_inputs_per_obj"][i]
                              for i in range(B))
        if not frame_has_input:
            output_dict = st["output_dict"]
            cfi = st["consolidated_frame_inds"]
            cfi["cond_frame_outputs"].discard(frame_idx)
            cfi["non_cond_frame_outputs"].discard(frame_idx)
            out = output_dict["cond_frame_outputs"].pop(frame_idx, None)
            if out is not None:
                output_dict["non_cond_frame_outputs"][frame_idx] = out
                st["frames_already_tracked"].pop(frame_idx, None)
            for i in range(B):
                od = st["output_dict_per_obj"][i]
                o = od["cond_frame_outputs"].pop(frame_idx, None)
                if o is not None:
                    od["non_cond_frame_outputs"][frame_idx] = o
            if len(output_dict["cond_frame_outputs"]) == 0:
                self._reset_tracking_results(st)
        if not need_output:
            return
        is_cond = any(frame_idx in t["cond_frame_outputs"] for t in temp.values())
        consolidated = self._consolidate_temp_output_across_obj(st, frame_idx, is_cond, run_mem_encoder=False,
                                                                consolidate_at_video_res=True)
        _, video_res_masks = self._get_orig_video_res_output(st, consolidated["pred_masks_video_res"])
        return frame_idx, st["obj_ids"], video_res_masks

    @torch.inference_mode()
    def reset_state(self, inference_state):
        """Remove all prompts and results (reference :848-860)."""
        st = inference_state
        self._reset_tracking_results(st)
        st["obj_id_to_idx"].clear()
        st["obj_idx_to_id"].clear()
        st["obj_ids"].clear()
        st["point_inputs_per_obj"].clear()
        st["mask_inputs_per_obj"].clear()
        st["output_dict_per_obj"].clear()
        st["temp_output_dict_per_obj"].clear()

    def _reset_tracking_results(self, st):
        for v in st["point_inputs_per_obj"].values():
            v.clear()
        for v in st["mask_inputs_per_obj"].values():
            v.clear()
        for v in st["output_dict_per_obj"].values():
            v["cond_frame_outputs"].clear()
            v["non_cond_frame_outputs"].clear()
        for v in st["temp_output_dict_per_obj"].values():
            v["cond_frame_outputs"].clear()
            v["non_cond_frame_outputs"].clear()
        st["output_dict"]["cond_frame_outputs"].clear()
        st["output_dict"]["non_cond_frame_outputs"].clear()
        st["consolidated_frame_inds"]["cond_frame_outputs"].clear()
        st["consolidated_frame_inds"]["non_cond_frame_outputs"].clear()
        st["tracking_has_started"] = False
        st["frames_already_tracked"].clear()

    @torch.inference_mode()
    def remove_object(self, inference_state, obj_id, strict=False, need_output=True):
        """Remove an object id from the tracking state (reference :1041-1150)."""
        st = inference_state
        old_idx = st["obj_id_to_idx"].get(obj_id, None)
        updated_frames = []
        if old_idx is None:
            if not strict:
                return st["obj_ids"], updated_frames
            raise RuntimeError(f"Cannot remove object id {obj_id} as it doesn't exist. "
                               f"All existing object ids: {st['obj_ids']}.")
        if len(st["obj_id_to_idx"]) == 1:
            self.reset_state(st)
            return st["obj_ids"], updated_frames
        input_frames = set(st["point_inputs_per_obj"][old_idx]) | set(st["mask_inputs_per_obj"][old_idx])
        for frame_idx in input_frames:
            self.clear_all_prompts_in_frame(st, frame_idx, obj_id, need_output=False)
        old_ids = st["obj_ids"]
        old_inds = list(range(len(old_ids)))
        remain = [i for i in old_inds if i != old_idx]
        new_ids = [old_ids[i] for i in remain]
        new_inds = list(range(len(new_ids)))
        old_to_new = dict(zip(remain, new_inds))
        st["obj_id_to_idx"] = dict(zip(new_ids, new_inds))
        st["obj_idx_to_id"] = dict(zip(new_inds, new_ids))
        st["obj_ids"] = new_ids

        def remap(container):
            kvs = []
            for k in old_inds:
                v = container.pop(k)
                if k in old_to_new:
                    kvs.append((old_to_new[k], v))
            container.update(kvs)

        for key in ("point_inputs_per_obj", "mask_inputs_per_obj", "output_dict_per_obj", "temp_output_dict_per_obj"):
            remap(st[key])
        if "frames_tracked_per_obj" in st:  # EfficientTAM's per-object bookkeeping
            remap(st["frames_tracked_per_obj"])
        if st.get("_store") is not None:
            st["_store"].select_objects(remain)
            self._refresh_views(st)
        if need_output:
            temp = st["temp_output_dict_per_obj"]
            for frame_idx in input_frames:
                is_cond = any(frame_idx in t["cond_frame_outputs"] for t in temp.values())
                consolidated = self._consolidate_temp_output_across_obj(st, frame_idx, is_cond, run_mem_encoder=False,
                                                                        consolidate_at_video_res=True)
                _, video_res_masks = self._get_orig_video_res_output(st, consolidated["pred_masks_video_res"])
                updated_frames.append((frame_idx, video_res_masks))
        return st["obj_ids"], updated_frames

    def _clear_non_cond_mem_around_input(self, st, frame_idx):
        """Drop non-conditioning memories around an edited frame (reference :1152-1172)."""
        r = self.memory_temporal_stride_for_eval
        lo, hi = frame_idx - r * self.num_maskmem, frame_idx + r * self.num_maskmem
        for t in range(lo, hi + 1):
            st["output_dict"]["non_cond_frame_outputs"].pop(t, None)
            for od in st["output_dict_per_obj"].values():
                od["non_cond_frame_outputs"].pop(t, None)


class _LockstepFeatures:
    """Backbone features of S sessions at one frame index for propagate_in_videos: each batched encoder pass covers
    `encoder_batch // S` (>= 1) consecutive time steps of all S videos, time-major, so the S frames of one time step are
    one contiguous [S, ...] block (what the batched frame graph reads)."""

    def __init__(self, pred, states, step):
        self.pred, self.states, self.step = pred, states, step
        self.S = len(states)
        self.nt = max(1, pred.encoder_batch // self.S)
        self.cache = {}

    def get(self, t):
        hit = self.cache.get(t)
        if hit is not None:
            return hit
        pred, states, S = self.pred, self.states, self.S
        T = states[0]["num_frames"]
        ts = [x for x in (t + i * self.step for i in range(self.nt)) if 0 <= x < T]
        n = len(ts) * S
        if pred.use_cuda_graphs and len(ts) == self.nt and n > 1:
            graph, static_in, out, n_kernels = pred._encoder_graph(n)
            for j, tt in enumerate(ts):
                for i, st in enumerate(states):
                    static_in[j * S + i].copy_(pred._frame(st, tt), non_blocking=True)
            graph.replay()
            pred._static_gen += 1  # the static outputs other sessions may have cached views of are overwritten
            _lib.launch_count += n_kernels
        else:
            imgs = torch.stack([pred._frame(st, tt) for tt in ts for st in states]).contiguous()
            out = pred.engine().encode_frames(imgs)
        self.cache = {tt: {k: v[j * S:(j + 1) * S] for k, v in out.items()} for j, tt in enumerate(ts)}
        return self.cache[t]


class SAM2VideoPredictorNPZ(SAM2VideoPredictor):
    """Variant whose `init_state` takes pre-normalised frames (sam2_video_predictor_npz.py:44-115)."""

    @torch.inference_mode()
    def init_state(self, images, video_height, video_width, offload_video_to_cpu=False, offload_state_to_cpu=False,
                   async_loading_frames=False):
        return self._new_state(images, video_height, video_width, offload_video_to_cpu, offload_state_to_cpu)


class _PerObjectTracked(dict):
    """`frames_tracked_per_obj` of the EfficientTAM session state (efficienttam_video_predictor.py:103): one dict per object
    index, created on first use."""

    def __missing__(self, key):
        self[key] = {}
        return self[key]


class EfficientTAMVideoPredictor(SAM2VideoPredictor):
    """efficient_track_anything/efficienttam_video_predictor.py: the same session API over the EfficientTAM model (ViT
    trunk + ViTDetNeck; no high-resolution decoder features, no pointer temporal encoding, no no_obj_embed_spatial).

    The reference keeps ALL state per object (conditioning frames, tracked frames) and runs the objects of a session one
    at a time (:489-628).  Objects are independent on this path, so while every object is prompted on the same frames the
    batched frame of the base class computes exactly the per-object results and is used.  As soon as the objects' prompt
    frames differ (or an object is added after tracking started, or a prompt is cleared) the session switches -- for
    good, until reset_state -- to the reference's per-object schedule: each object is tracked alone (one-object frame
    graph addressed at its column of the frame store) against its own conditioning / non-conditioning outputs, so an
    object without input on a frame another object is prompted on is tracked normally there."""
    _config_base = EtamTiConfig
    _abi = staticmethod(synth.etam_state_dict_abi)

    def _new_state(self, *args, **kwargs):
        st = super()._new_state(*args, **kwargs)
        st["frames_tracked_per_obj"] = _PerObjectTracked()
        return st

    def _obj_id_to_idx(self, st, obj_id):
        """New objects are always allowed, also after tracking started (reference :127-159)."""
        idx = st["obj_id_to_idx"].get(obj_id, None)
        if idx is None and st["tracking_has_started"]:
            started, st["tracking_has_started"] = True, False
            try:
                idx = super()._obj_id_to_idx(st, obj_id)
            finally:
                st["tracking_has_started"] = started
            st["_per_object"] = True
        elif idx is None:
            idx = super()._obj_id_to_idx(st, obj_id)
        st["frames_tracked_per_obj"][idx]  # noqa: B018  (creates the object's entry)
        return idx

    def _tracked_info(self, st, obj_idx, frame_idx):
        return st["frames_tracked_per_obj"][obj_idx].get(frame_idx)  # per object (reference :238-247)

    def _reset_tracking_results(self, st):
        super()._reset_tracking_results(st)
        st["frames_tracked_per_obj"].clear()
        st.pop("_per_object", None)

    # ------------------------------------------------------------------ per-object schedule
    def _per_object_mode(self, st):
        if not st.get("_per_object"):
            inputs = [set(st["point_inputs_per_obj"][i]) | set(st["mask_inputs_per_obj"][i])
                      for i in range(self._get_obj_num(st))]
            if any(s != inputs[0] for s in inputs[1:]):
                st["_per_object"] = True
        return bool(st.get("_per_object"))

    def _cond_frames(self, st):
        if not st.get("_per_object"):
            return super()._cond_frames(st)
        return {t for od in st["output_dict_per_obj"].values() for t in od["cond_frame_outputs"]}

    def _frames_without_tracking(self, st):
        if not st.get("_per_object"):
            return super()._frames_without_tracking(st)
        ods = list(st["output_dict_per_obj"].values())
        return set.intersection(*[set(od["cond_frame_outputs"]) for od in ods]) if ods else set()

    def _clear_obj_non_cond_mem_around_input(self, st, frame_idx, obj_idx):
        r = self.memory_temporal_stride_for_eval
        non_cond = st["output_dict_per_obj"][obj_idx]["non_cond_frame_outputs"]
        for t in range(frame_idx - r * self.num_maskmem, frame_idx + r * self.num_maskmem + 1):
            non_cond.pop(t, None)

    @torch.inference_mode()
    def propagate_in_video_preflight(self, inference_state):
        """Per-object consolidation of the temporary outputs (reference :489-552)."""
        st = inference_state
        if not self._per_object_mode(st):
            return super().propagate_in_video_preflight(st)
        self._sync_engine()
        B = self._get_obj_num(st)
        if B == 0:
            raise RuntimeError("No input points or masks are provided for any object; please add inputs first.")
        st["tracking_has_started"] = True
        store = self._ensure_store(st)
        for i in range(B):
            od, tmp = st["output_dict_per_obj"][i], st["temp_output_dict_per_obj"][i]
            for storage_key in ("non_cond_frame_outputs", "cond_frame_outputs"):
                for t, out in list(tmp[storage_key].items()):
                    mem_tok = out.get("_mem_tok")
                    if mem_tok is None:
                        mem_tok = self._run_memory_encoder(st, t, 1, out["pred_masks"], out["object_score_logits"], True)
                    store.mem[t, i].copy_(mem_tok[0])
                    store.ptr[t, i].copy_(out["obj_ptr"][0])
                    store.score[t, i].copy_(out["object_score_logits"][0])
                    store.masks[t, i].copy_(out["pred_masks"][0])
                    od[storage_key][t] = self._obj_slot_views(st, t, i)
                    if self.clear_non_cond_mem_around_input:
                        self._clear_obj_non_cond_mem_around_input(st, t, i)
                tmp[storage_key].clear()
            if len(od["cond_frame_outputs"]) == 0:
                raise RuntimeError(f"No input points or masks are provided for object id {self._obj_idx_to_id(st, i)}; "
                                   "please add inputs first.")
            for t in od["cond_frame_outputs"]:
                od["non_cond_frame_outputs"].pop(t, None)

    def _propagate_loop(self, st, order, reverse, clear_non_cond_mem, B):
        if not st.get("_per_object"):
            yield from super()._propagate_loop(st, order, reverse, clear_non_cond_mem, B)
            return
        store, obj_ids = st["_store"], st["obj_ids"]
        for frame_idx in _progress(order, "propagate in video"):
            for i in range(B):  # one object at a time against its own outputs (reference :592-628)
                od = st["output_dict_per_obj"][i]
                if frame_idx in od["cond_frame_outputs"]:
                    if self.clear_non_cond_mem_around_input:
                        self._clear_obj_non_cond_mem_around_input(st, frame_idx, i)
                else:
                    out, _ = self._track_frame(st, frame_idx, 1, reverse, od, obj0=i)
                    od["non_cond_frame_outputs"][frame_idx] = out
                st["frames_tracked_per_obj"][i][frame_idx] = {"reverse": reverse}
            st["frames_already_tracked"][frame_idx] = {"reverse": reverse}
            _, video_res_masks = self._get_orig_video_res_output(st, store.masks[frame_idx])
            yield frame_idx, obj_ids, video_res_masks

    @torch.inference_mode()
    def clear_all_prompts_in_frame(self, inference_state, frame_idx, obj_id, need_output=True):
        """Per-object removal of a frame's prompts (reference :642-682): the object's conditioning output on that frame is
        downgraded to a non-conditioning one; nothing is reset.  The prompt sets may differ afterwards, so the session
        continues on the per-object schedule."""
        st = inference_state
        obj_idx = self._obj_id_to_idx(st, obj_id)
        st["_per_object"] = True
        st["point_inputs_per_obj"][obj_idx].pop(frame_idx, None)
        st["mask_inputs_per_obj"][obj_idx].pop(frame_idx, None)
        temp = st["temp_output_dict_per_obj"]
        temp[obj_idx]["cond_frame_outputs"].pop(frame_idx, None)
        temp[obj_idx]["non_cond_frame_outputs"].pop(frame_idx, None)
        od = st["output_dict_per_obj"][obj_idx]
        out = od["cond_frame_outputs"].pop(frame_idx, None)
        if out is not None:
            od["non_cond_frame_outputs"][frame_idx] = out
            st["frames_tracked_per_obj"][obj_idx].pop(frame_idx, None)
        if not need_output:
            return
        is_cond = any(frame_idx in t["cond_frame_outputs"] for t in temp.values())
        consolidated = self._consolidate_temp_output_across_obj(st, frame_idx, is_cond, run_mem_encoder=False,
                                                                consolidate_at_video_res=True)
        _, video_res_masks = self._get_orig_video_res_output(st, consolidated["pred_masks_video_res"])
        return frame_idx, st["obj_ids"], video_res_masks


class EfficientTAMVideoPredictorNPZ(EfficientTAMVideoPredictor):
    """efficient_track_anything/efficienttam_video_predictor_npz.py: init_state from an in-memory image tensor."""

    @torch.inference_mode()
    def init_state(self, images, video_height, video_width, offload_video_to_cpu=False, offload_state_to_cpu=False,
                   async_loading_frames=False):
        return self._new_state(images, video_height, video_width, offload_video_to_cpu, offload_state_to_cpu)


def _select_closest_cond_frames(frame_idx, cond_frame_outputs, max_cond_frame_num):
    """Conditioning frames closest in time (sam2_utils.py:19-61)."""
    if max_cond_frame_num == -1 or len(cond_frame_outputs) <= max_cond_frame_num:
        return cond_frame_outputs, {}
    assert max_cond_frame_num >= 2, "we should allow using 2+ conditioning frames"
    chosen = {}
    before = max((t for t in cond_frame_outputs if t < frame_idx), default=None)
    if before is not None:
        chosen[before] = cond_frame_outputs[before]
    after = min((t for t in cond_frame_outputs if t >= frame_idx), default=None)
    if after is not None:
        chosen[after] = cond_frame_outputs[after]
    rest = sorted((t for t in cond_frame_outputs if t not in chosen), key=lambda t: abs(t - frame_idx))
    for t in rest[: max_cond_frame_num - len(chosen)]:
        chosen[t] = cond_frame_outputs[t]
    return chosen, {t: v for t, v in cond_frame_outputs.items() if t not in chosen}
