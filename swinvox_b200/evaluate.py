"""Batched evaluation driver: the caller of the hot path, replacing the reference's batch-1 loop with
5-6 host synchronisations per threshold per sample (core/test.py:107-206).

* batches of B objects go through the pipeline; the host-to-device copy of batch i+1 runs on a copy stream while
  batch i computes (two pinned staging slots, one device staging buffer per slot);
* IoU / F-score are computed from the integer counters of the metric kernel ON THE DEVICE and accumulated per
  taxonomy with index_add; the BCE losses come from the same kernel; nothing is read back until `finish()`;
* `finish()` does one device-to-host copy and prints the reference's two tables (same text, core/test.py:222-262)
  and returns the reference's return value (max over thresholds of the sample-weighted mean IoU).
"""
import io

import numpy as np
import torch

from .metrics import VoxelMetrics


class BatchedEvaluator:
    def __init__(self, recon, batch_size, n_views, taxonomies=None, on_batch=None):
        """recon: pipeline.Reconstructor; taxonomies: {taxonomy_id: {"taxonomy_name": ..., "baseline": {...}}} as loaded
        from the dataset's taxonomy file (core/test.py:38-41); ids may be any hashable.
        on_batch(logits_host [n,32,32,32], counts_host [n,T,5]): optional consumer of every batch's voxels; the
        device-to-host copy runs on its own stream into two pinned slots and the callback fires once it has landed
        (up to two batches later, or in finish())."""
        self.recon, self.B, self.V = recon, int(batch_size), int(n_views)
        self.cfg = recon.cfg
        self.thresholds = [float(t) for t in self.cfg.TEST.VOXEL_THRESH]
        self.taxonomies = taxonomies or {}
        self.dev = recon.device
        self.cuda = self.dev.type == "cuda"
        self.inbuf = recon.input_buffer(self.B, self.V)
        self.gt = torch.zeros(self.B, 32, 32, 32, device=self.dev)
        T = len(self.thresholds)
        self._tax_index = {}
        self._cap = 64
        self._iou = torch.zeros(self._cap, T, dtype=torch.float64, device=self.dev)
        self._fsc = torch.zeros(self._cap, T, dtype=torch.float64, device=self.dev)
        self._cnt = torch.zeros(self._cap, dtype=torch.float64, device=self.dev)
        self._loss = torch.zeros(2, dtype=torch.float64, device=self.dev)   # sums of EDLoss, RLoss over samples
        self.n_samples = 0
        self.metrics_ref = VoxelMetrics(self.thresholds)
        self.metrics_mrg = VoxelMetrics(self.thresholds)
        if self.cuda:
            self.copy_stream = torch.cuda.Stream(self.dev)
            self.stage = [torch.empty_like(self.inbuf) for _ in range(2)]
            self.stage_gt = [torch.empty_like(self.gt) for _ in range(2)]
            self.stage_tix_host = [torch.zeros(self.B, dtype=torch.int64).pin_memory() for _ in range(2)]
            self.stage_tix = [torch.zeros(self.B, dtype=torch.int64, device=self.dev) for _ in range(2)]
            self.tix = torch.zeros(self.B, dtype=torch.int64, device=self.dev)
            self.ready = [torch.cuda.Event() for _ in range(2)]
            self.consumed = [torch.cuda.Event() for _ in range(2)]
            for e in self.consumed:
                e.record()
            self.on_batch = on_batch
            if on_batch is not None:
                T5 = (self.B, T, 5)
                self.out_stream = torch.cuda.Stream(self.dev)
                self.out_logits = [torch.empty(self.B, 32, 32, 32).pin_memory() for _ in range(2)]
                self.out_counts = [torch.empty(T5, dtype=torch.int32).pin_memory() for _ in range(2)]
                self.out_done = [torch.cuda.Event() for _ in range(2)]
                self.out_n = [0, 0]
                self.out_last = None   # out_done event of the newest D2H copy: it reads plan-owned buffers
        else:
            self.on_batch = on_batch
        self._slot = 0
        self._pending = None   # (slot, taxonomy index tensor, n_valid)

    # ---- staging --------------------------------------------------------------------------------
    def _tax_ids(self, taxonomy_ids):
        idx = []
        for t in taxonomy_ids:
            if t not in self._tax_index:
                self._tax_index[t] = len(self._tax_index)
                if len(self._tax_index) > self._cap:
                    raise ValueError("more than %d taxonomies" % self._cap)
            idx.append(self._tax_index[t])
        return idx

    def submit(self, taxonomy_ids, images, gt):
        """images [n,V,3,224,224], gt [n,32,32,32] host tensors (pinned for a truly asynchronous copy), n <= B;
        a short last batch is padded by repeating its final object and the padding is masked out of the sums."""
        n = images.shape[0]
        if n > self.B or tuple(images.shape[1:]) != (self.V, 3, 224, 224):
            raise ValueError(f"expected at most {self.B} objects of shape [{self.V},3,224,224], got {tuple(images.shape)}")
        tix = self._tax_ids(taxonomy_ids)
        tix = tix + [tix[-1]] * (self.B - n)
        s = self._slot
        self._slot ^= 1
        if self.cuda:
            self.consumed[s].synchronize()   # (host) the slot's pinned index buffer was read by its previous batch
            self.stage_tix_host[s].copy_(torch.tensor(tix, dtype=torch.int64))
            with torch.cuda.stream(self.copy_stream):
                self.consumed[s].wait(self.copy_stream)     # the compute stream has read this slot's previous batch
                self.stage_tix[s].copy_(self.stage_tix_host[s], non_blocking=True)
                self.stage[s][:n].copy_(images, non_blocking=True)
                self.stage_gt[s][:n].copy_(gt, non_blocking=True)
                if n < self.B:
                    self.stage[s][n:] = self.stage[s][n - 1]
                    self.stage_gt[s][n:] = self.stage_gt[s][n - 1]
                self.ready[s].record(self.copy_stream)
        prev, self._pending = self._pending, (s, tix, n, images if not self.cuda else None, gt if not self.cuda else None)
        if prev is not None:
            self._run(prev)

    def _run(self, pending):
        s, tix, n, images, gt = pending
        if self.cuda:
            cur = torch.cuda.current_stream(self.dev)
            self.ready[s].wait(cur)
            if self.on_batch is not None and self.out_last is not None:
                # the previous batch's device-to-host copies read the plan-owned logits / counters this run overwrites
                self.out_last.wait(cur)
            self.inbuf.copy_(self.stage[s], non_blocking=True)      # device-to-device: the plans read fixed addresses
            self.gt.copy_(self.stage_gt[s], non_blocking=True)
            self.tix.copy_(self.stage_tix[s], non_blocking=True)
            self.consumed[s].record(cur)
        else:   # the CPU twin of the tests
            self.inbuf[:n].copy_(images)
            self.gt[:n].copy_(gt)
            if n < self.B:
                self.inbuf[n:] = self.inbuf[n - 1]
                self.gt[n:] = self.gt[n - 1]
        with torch.no_grad():
            r = self.recon   # core/test.py:120-130, incl. the epoch gates of a loaded checkpoint
            raw, gen = r.decoder(r.encoder(self.inbuf))
            if r.merger is not None and r._gate("EPOCH_START_USE_MERGER"):
                merged = r.merger(raw, gen)
            else:
                merged = r.view_mean(gen)
            refined = r.refiner(merged) if (r.refiner is not None and r._gate("EPOCH_START_USE_REFINER")) else merged
        counts, bce_r = self.metrics_ref.counts_and_bce(refined, self.gt)
        if refined is merged:
            bce_m = bce_r
        else:
            _, bce_m = self.metrics_mrg.counts_and_bce(merged, self.gt)
        iou, fsc = VoxelMetrics.scores(counts)
        valid = torch.zeros(self.B, dtype=torch.float64, device=self.dev)
        valid[:n] = 1.0
        t = self.tix if self.cuda else torch.tensor(tix, device=self.dev)
        self._iou.index_add_(0, t, iou.to(torch.float64) * valid[:, None])
        self._fsc.index_add_(0, t, fsc.to(torch.float64) * valid[:, None])
        self._cnt.index_add_(0, t, valid)
        self._loss += torch.stack([(bce_m * valid).sum(), (bce_r * valid).sum()]) * 10.0
        self.n_samples += n
        self.last_logits = refined   # plan-owned: valid until the next batch runs (clone to keep)
        if self.on_batch is not None:
            if self.cuda:
                self._deliver(s)                                   # the slot's previous occupant, if still undelivered
                done = torch.cuda.Event()
                done.record(torch.cuda.current_stream(self.dev))
                with torch.cuda.stream(self.out_stream):
                    done.wait(self.out_stream)
                    self.out_logits[s].copy_(refined, non_blocking=True)
                    self.out_counts[s].copy_(counts, non_blocking=True)
                    self.out_done[s].record(self.out_stream)
                self.out_last = self.out_done[s]
                self.out_n[s] = n
            else:
                self.on_batch(refined[:n].clone(), counts[:n].clone())

    def _deliver(self, s):
        if self.out_n[s]:
            self.out_done[s].synchronize()
            n, self.out_n[s] = self.out_n[s], 0
            self.on_batch(self.out_logits[s][:n], self.out_counts[s][:n])

    # ---- results --------------------------------------------------------------------------------
    def finish(self, print_tables=True, file=None):
        """runs the batch still in flight, reads the accumulators back (ONE device-to-host copy) and returns
        (max_iou, report dict); prints the reference's tables"""
        if self._pending is not None:
            self._run(self._pending)
            self._pending = None
        if self.on_batch is not None and self.cuda:
            for s in (self._slot, self._slot ^ 1):     # oldest first
                self._deliver(s)
        mer = getattr(self.recon, "merger", None)
        if mer is not None and hasattr(mer, "saturated") and mer.saturated():
            import warnings
            warnings.warn("swinvox_b200: merger activations exceeded fp16's range (65504) during this evaluation; the fp16 "
                          "tensor-core operands saturated.  The merger now uses tf32 operands: re-run the evaluation.")
        enc = getattr(self.recon, "encoder", None)
        if enc is not None and hasattr(enc, "attention_saturated") and enc.attention_saturated():
            import warnings
            warnings.warn("swinvox_b200: a window-attention value exceeded fp16's range (65504) during this evaluation; the "
                          "fp16 P.V operands saturated and the affected outputs are not within the parity tolerance.")
        self._reduce_over_ranks()
        nt = len(self._tax_index)
        packed = torch.cat([self._iou[:nt].flatten(), self._fsc[:nt].flatten(), self._cnt[:nt], self._loss]).cpu().numpy()
        T = len(self.thresholds)
        iou_sum = packed[:nt * T].reshape(nt, T)
        fsc_sum = packed[nt * T:2 * nt * T].reshape(nt, T)
        cnt = packed[2 * nt * T:2 * nt * T + nt]
        loss = packed[-2:]
        test_iou, test_fscore = {}, {}
        for tid, k in self._tax_index.items():
            test_iou[tid] = {"n_samples": int(cnt[k]), "iou": iou_sum[k] / cnt[k]}
            test_fscore[tid] = {"n_samples": int(cnt[k]), "fscore": fsc_sum[k] / cnt[k]}
        n = max(self.n_samples, 1)
        mean_iou, mean_fscore = iou_sum.sum(0) / n, fsc_sum.sum(0) / n
        report = {"test_iou": test_iou, "test_fscore": test_fscore, "mean_iou": mean_iou, "mean_fscore": mean_fscore,
                  "encoder_loss": loss[0] / n, "refiner_loss": loss[1] / n, "n_samples": self.n_samples}
        if print_tables:
            print(self.tables(report), end="", file=file)
        return float(np.max(mean_iou)) if nt else 0.0, report

    def _reduce_over_ranks(self):
        """data-parallel evaluation (one process per GPU, each rank fed its own shard of the test set): the per-taxonomy
        sums are added over the ranks -- the only exchange, one small all_reduce (NCCL on GPUs) -- so every rank
        reports the global result.  Taxonomies are re-indexed in a rank-independent order first."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return
        keys = [None] * dist.get_world_size()
        dist.all_gather_object(keys, list(self._tax_index))
        order = []
        for ks in keys:                      # rank 0's first-seen order, then what later ranks add
            for k in ks:
                if k not in order:
                    order.append(k)
        if len(order) > self._cap:
            raise ValueError("more than %d taxonomies" % self._cap)
        perm = torch.tensor([self._tax_index.get(k, -1) for k in order], device=self.dev)
        have = (perm >= 0)
        src = perm.clamp_min(0)

        def remap(t):
            out = torch.zeros_like(t)
            sel = t[src] * (have.to(t.dtype).view(-1, *([1] * (t.dim() - 1))))
            out[:len(order)] = sel
            return out

        self._iou, self._fsc, self._cnt = remap(self._iou), remap(self._fsc), remap(self._cnt)
        self._tax_index = {k: i for i, k in enumerate(order)}
        n = torch.tensor([float(self.n_samples)], dtype=torch.float64, device=self.dev)
        for t in (self._iou, self._fsc, self._cnt, self._loss, n):
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        self.n_samples = int(round(n.item()))

    def tables(self, report):
        """the text of core/test.py:222-262"""
        out = io.StringIO()
        names = {tid: self.taxonomies.get(tid, {}).get("taxonomy_name", str(tid)) for tid in report["test_iou"]}

        def header(title):
            out.write(title + "\n")
            out.write("Taxonomy\t#Sample\tBaseline\t" + "".join(f"t={th:.2f}\t" for th in self.thresholds) + "\n")

        header('============================ TEST RESULTS (IoU) ============================')
        for tid, rec in report["test_iou"].items():
            base = self.taxonomies.get(tid, {}).get("baseline", {}).get(f"{self.V}-view")
            out.write(f"{names[tid].ljust(8)}\t{rec['n_samples']}\t" + (f"{base:.4f}\t\t" if base is not None else "N/a\t\t"))
            out.write("".join(f"{v:.4f}\t" for v in rec["iou"]) + "\n")
        out.write("Overall \t\t\t\t" + "".join(f"{v:.4f}\t" for v in report["mean_iou"]) + "\n\n")
        header('========================== TEST RESULTS (F-score) ==========================')
        for tid, rec in report["test_fscore"].items():
            out.write(f"{names[tid].ljust(8)}\t{rec['n_samples']}\tN/a\t\t")
            out.write("".join(f"{v:.4f}\t" for v in rec["fscore"]) + "\n")
        out.write("Overall \t\t\t\t" + "".join(f"{v:.4f}\t" for v in report["mean_fscore"]) + "\n\n")
        return out.getvalue()
