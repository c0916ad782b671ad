"""Frame-parallel image encoder running AHEAD of the sequential propagation.

The Hiera encoder is frame-independent while the memory-attention propagation is strictly sequential in t
(reference: sam2_video_predictor.py:879-910 encodes a frame lazily right before it is tracked; training/model/sam2.py:107-110
encodes all frames up front).  Here the frames of the tracking order are cut into batches of `encoder_batch` frames
(`BatchPlan`) that a *producer* encodes while the consumer tracks the previous batch:

  * `PartitionProducer` -- same GPU: the encoder replays its CUDA graph on a stream of a CUDA green context that owns a
    fixed subset of the SMs (`SmPartition`), so the throughput-bound encoder and the latency-bound tracked frame (8-144
    CTAs per kernel) run side by side instead of back to back;
  * `RemoteProducer` / `serve_clip_encoder` -- single long clip on several GPUs: the other ranks encode their share of
    the batches and ship the three FPN levels to the propagation rank with point-to-point NCCL sends over NVLink
    (no collective anywhere: the path has no reduction step).

`FeaturePipeline` is the consumer-side bookkeeping shared by both (which batch is in flight in which slot, when a slot may
be overwritten); it only needs `launch` / `wait` from a producer, so the protocol is tested on CPU with gloo and a stand-in
encoder (tests/test_host_logic.py).
"""
import torch
import torch.distributed as dist

BF16, F32 = torch.bfloat16, torch.float32

# what the encoder leaves behind per frame (Engine.encode_frames): name -> (shape, dtype)
def feature_specs(feat=32):
    """`feat`: side of the stride-16 feature map (32 at 512^2, 64 for Hiera-B+ at 1024^2)."""
    T = feat * feat
    return (("feat", (T, 256), F32), ("feat_bf16", (T, 256), BF16), ("feat_s1", (4 * T, 64), F32),
            ("feat_s0", (16 * T, 32), F32))


FEATURE_SPECS = feature_specs(32)
# what travels between GPUs: 4 MiB per frame.  The bf16 copy of `feat` is re-derived on the receiving side (one cast
# kernel); the three levels stay fp32 on the wire -- the decoder reads feat_s0 / feat_s1 as fp32 operands and the Dice bar
# leaves no room for another rounding, while 4 MiB x ~1400 frames/s is < 1 % of one NVLink 5 direction.
WIRE_SPECS = tuple(spec for spec in FEATURE_SPECS if spec[0] != "feat_bf16")


class BatchPlan:
    """Frames first, first+step, ..., last (tracking order) cut into batches of n.  With include_tail=False a final
    partial batch is left to the caller (the captured encoder graph has a fixed batch size)."""

    def __init__(self, first, last, step, n, include_tail=True):
        if step not in (1, -1) or n < 1:
            raise ValueError("step must be +1 / -1 and n >= 1")
        self.first, self.last, self.step, self.n = int(first), int(last), int(step), int(n)
        count = (self.last - self.first) * self.step + 1
        self.count = max(0, count)
        full, rem = divmod(self.count, self.n)
        self.num_batches = full + (1 if rem and include_tail else 0)
        self.covered = min(self.count, self.num_batches * self.n)

    def frames(self, j):
        lo = j * self.n
        hi = min(lo + self.n, self.covered)
        return [self.first + i * self.step for i in range(lo, hi)]

    def batch_of(self, t):
        i = (t - self.first) * self.step
        return (i // self.n, i % self.n) if 0 <= i < self.covered else None

    def header(self):
        return [self.first, self.last, self.step, self.n]


def alloc_feature_slot(n, device, feat=32):
    return {name: torch.empty((n,) + shape, dtype=dt, device=device) for name, shape, dt in feature_specs(feat)}


class FeaturePipeline:
    """Consumer side: keeps `depth` batches in flight ahead of the one being tracked, in depth + 1 slots.  A slot is
    handed back to the producer only when the consumer has asked for the first frame of the *next* batch, i.e. after every
    use of the slot's previous contents has been enqueued on the consumer's stream."""

    def __init__(self, plan, producer, depth=1):
        self.plan, self.producer, self.depth = plan, producer, max(1, int(depth))
        self.nslots = self.depth + 1
        self.launched = 0   # next batch to hand to the producer
        self.installed = 0  # next batch the consumer will take
        self.current = None  # (batch index, feature dict) of the batch being consumed

    def _launch_upto(self, bound):
        while self.launched < min(bound, self.plan.num_batches):
            k = self.launched
            self.producer.launch(k, self.plan.frames(k), k % self.nslots)
            self.launched += 1

    def get(self, t):
        """Features of frame t ({name: tensor}) or None if t is not (or no longer) served by the pipeline."""
        where = self.plan.batch_of(t)
        if where is None:
            return None
        j, pos = where
        if self.current is not None and j == self.current[0]:
            return {k: v[pos] for k, v in self.current[1].items()}
        if j < self.installed:
            return None  # an older batch: its slot may already be overwritten
        while self.installed <= j:
            k = self.installed
            self._launch_upto(k + 1)
            feats = self.producer.wait(k)
            self.installed += 1
            self.current = (k, feats)
            self._launch_upto(self.installed + self.depth)
        return {k: v[pos] for k, v in self.current[1].items()}

    def close(self):
        """Take delivery of whatever is still owed (a producer on another rank sends every batch of the plan)."""
        if getattr(self.producer, "must_drain", False):
            while self.installed < self.plan.num_batches:
                k = self.installed
                self._launch_upto(k + 1)
                self.producer.wait(k)
                self.installed += 1
        self.current = None


# ------------------------------------------------------------------------------------------------
# same GPU: encoder on an SM partition
# ------------------------------------------------------------------------------------------------
class SmPartition:
    """A stream whose kernels run on a fixed subset of the SMs (CUDA green context, driver API through cuda-python).
    `sms` is the size actually granted (the driver rounds the request up to its partition granularity)."""

    def __init__(self, device, sms):
        from cuda.bindings import driver as drv

        def ck(res):
            if res[0] != drv.CUresult.CUDA_SUCCESS:
                raise RuntimeError(f"CUDA driver call failed: {res[0]}")
            return res[1] if len(res) == 2 else res[1:]

        device = torch.device(device)
        torch.zeros(1, device=device)  # the primary context must exist before a green context is derived from it
        ck(drv.cuInit(0))
        dev = ck(drv.cuDeviceGet(device.index if device.index is not None else torch.cuda.current_device()))
        full = ck(drv.cuDeviceGetDevResource(dev, drv.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
        self.total_sms = int(full.sm.smCount)
        groups, n_groups, remaining = ck(drv.cuDevSmResourceSplitByCount(1, full, 0, int(sms)))
        if int(n_groups) < 1:
            raise RuntimeError("SM split produced no group")
        group = groups[0]
        self.sms = int(group.sm.smCount)
        if not (0 < self.sms < self.total_sms):
            raise RuntimeError(f"unusable SM partition: {self.sms} of {self.total_sms}")
        desc = ck(drv.cuDevResourceGenerateDesc([group], 1))
        self._ctx = ck(drv.cuGreenCtxCreate(desc, dev, drv.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
        self._raw = ck(drv.cuGreenCtxStreamCreate(self._ctx, drv.CUstream_flags.CU_STREAM_NON_BLOCKING, 0))
        self.stream = torch.cuda.ExternalStream(int(self._raw), device=device)
        # the complement (every SM not in the group) as a second green context
        self.rest_sms = int(remaining.sm.smCount)
        self.rest_stream = None
        if self.rest_sms > 0:
            rdesc = ck(drv.cuDevResourceGenerateDesc([remaining], 1))
            self._rest_ctx = ck(drv.cuGreenCtxCreate(rdesc, dev, drv.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
            self._rest_raw = ck(drv.cuGreenCtxStreamCreate(self._rest_ctx, drv.CUstream_flags.CU_STREAM_NON_BLOCKING, 0))
            self.rest_stream = torch.cuda.ExternalStream(int(self._rest_raw), device=device)
            self._rest_raw2 = ck(drv.cuGreenCtxStreamCreate(self._rest_ctx, drv.CUstream_flags.CU_STREAM_NON_BLOCKING, 0))
            self.rest_stream2 = torch.cuda.ExternalStream(int(self._rest_raw2), device=device)


class PartitionProducer:
    """Replays the captured image-encoder graph of slot (k mod 2) on the partition's stream.  The very first batch has
    nothing to overlap with (the consumer is waiting for it), so it runs on the consumer's stream with the whole device."""
    must_drain = False

    def __init__(self, predictor, st, partition, n):
        self.pred, self.st, self.part, self.n = predictor, st, partition, n
        self.pending = {}

    def launch(self, k, frames, slot):
        from . import _lib

        if k == 0:
            graph, static_in, out, n_kernels = self.pred._encoder_graph(self.n)  # full-device graph
            self.pred._load_frames(self.st, frames, static_in)
            graph.replay()
            self.pred._static_gen += 1  # (cached views of these static outputs held by any session are now stale)
            _lib.launch_count += n_kernels
            # its static outputs are shared with every other user of that graph: this batch gets its own copy (72 MB)
            self.pending[k] = (None, {name: t.clone() for name, t in out.items()})
            return
        graph, static_in, out, n_kernels = self.pred._encoder_graph(self.n, slot, self.part)
        enc = self.part.stream
        free = torch.cuda.Event()
        free.record()  # consumer stream: every use of this slot's previous contents is enqueued before this point
        enc.wait_event(free)
        with torch.cuda.stream(enc):
            self.pred._load_frames(self.st, frames, static_in)
            graph.replay()
            done = torch.cuda.Event()
            done.record()
        _lib.launch_count += n_kernels
        self.pending[k] = (done, out)

    def wait(self, k):
        done, out = self.pending.pop(k)
        if done is not None:
            torch.cuda.current_stream().wait_event(done)
        return out


# ------------------------------------------------------------------------------------------------
# several GPUs, one clip: encoder ranks ship FPN features to the propagation rank
# ------------------------------------------------------------------------------------------------
class RemoteEncoders:
    """Handle held by the propagation rank: which ranks encode, and the control channel that tells them what to encode
    (a 4-integer header broadcast per propagation pass; n = 0 ends their service loop)."""

    def __init__(self, encoder_ranks, device, group=None, src=None):
        self.ranks = list(encoder_ranks)
        if not self.ranks:
            raise ValueError("need at least one encoder rank")
        self.device, self.group = device, group
        self.src = dist.get_rank() if src is None else src

    def announce(self, plan):
        hdr = torch.tensor(plan.header(), dtype=torch.int64, device=self.device)
        dist.broadcast(hdr, src=self.src, group=self.group)

    def shutdown(self):
        hdr = torch.zeros(4, dtype=torch.int64, device=self.device)
        dist.broadcast(hdr, src=self.src, group=self.group)


class RemoteProducer:
    """Posts the receives of batch k (from encoder rank k mod E) into slot buffers; NCCL / gloo match the sends of a rank
    in the order they were issued, and every rank walks the plan in increasing k."""
    must_drain = True

    def __init__(self, remote, n, device, feat=32):
        self.remote, self.n = remote, n
        self.slots = [alloc_feature_slot(n, device, feat) for _ in range(len(remote.ranks) + 1)]
        self.pending = {}

    def launch(self, k, frames, slot):
        src = self.remote.ranks[k % len(self.remote.ranks)]
        bufs = {name: t[: len(frames)] for name, t in self.slots[slot].items()}
        ops_ = [dist.P2POp(dist.irecv, bufs[name], src, self.remote.group) for name, _, _ in WIRE_SPECS]
        self.pending[k] = (dist.batch_isend_irecv(ops_), bufs)

    def wait(self, k):
        works, bufs = self.pending.pop(k)
        for w in works:
            w.wait()
        bufs["feat_bf16"].copy_(bufs["feat"])  # the bf16 operand copy of `feat` is not sent: one cast on receipt
        return bufs


def serve_clip_encoder(encode, my_index, num_encoders, device, dst=0, group=None, on_plan=None, max_plans=None):
    """Service loop of an encoder rank.  `encode(frames, slot) -> {name: tensor [len(frames), ...]}` produces the features
    of the listed frame indices into buffers it owns per slot (0 / 1); they are read by the send until that send has
    completed, so a slot is reused only after waiting for its previous send.  Serves announced plans until the
    shutdown header arrives (or `max_plans` plans have been served); returns the number of frames encoded."""
    total, served = 0, 0
    while max_plans is None or served < max_plans:
        hdr = torch.zeros(4, dtype=torch.int64, device=device)
        dist.broadcast(hdr, src=dst, group=group)
        first, last, step, n = (int(x) for x in hdr.tolist())
        if n == 0:
            return total
        plan = BatchPlan(first, last, step, n, include_tail=True)
        if on_plan is not None:
            on_plan(plan)
        in_flight = {}
        mine = 0
        for k in range(plan.num_batches):
            if k % num_encoders != my_index:
                continue
            slot = mine % 2
            mine += 1
            for w in in_flight.pop(slot, ()):
                w.wait()
            frames = plan.frames(k)
            out = encode(frames, slot)
            ops_ = [dist.P2POp(dist.isend, out[name][: len(frames)], dst, group) for name, _, _ in WIRE_SPECS]
            in_flight[slot] = dist.batch_isend_irecv(ops_)
            total += len(frames)
        for works in in_flight.values():
            for w in works:
                w.wait()
        served += 1
    return total
