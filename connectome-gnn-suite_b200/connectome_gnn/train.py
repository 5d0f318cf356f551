"""Training / evaluation driver (host-side mirror of reference ``connectome_gnn/train.py``).

Same class, constructor, method names and return values as the reference ``Trainer``
(``train.py:19-127``).  Differences that do not change results:

* the loss is the fused ``cgnn_ce_fwd`` / ``cgnn_ce_bwd`` pair (``CrossEntropyLoss`` below keeps the
  ``loss_fn`` attribute of reference ``train.py:39``);
* per-batch losses / correct counts stay on the device and are read back ONCE per epoch; the
  float64 accumulation ``sum(loss_i * B_i) / sum(B_i)`` (``train.py:52-54,67-71``) is then done
  on the host over exactly the same fp32 per-batch values, so the returned numbers are the same
  while the per-batch device->host syncs of ``train.py:52,66,68`` disappear;
* under ``torch.distributed`` (one process per GPU; the loader shards every global batch) the loss
  is scaled by ``1 / B_global`` and parameter gradients are summed across ranks in one flat
  all-reduce, which reproduces the single-process ``CrossEntropyLoss`` mean.

There is no CPU compute path: ``device="cpu"`` (the reference default) is accepted for signature
compatibility but the model and batches run on the current CUDA device.
"""

from __future__ import annotations

import torch
import torch.nn as nn

from . import _engine

__all__ = ["Trainer", "CrossEntropyLoss", "FlatParams", "GraphedStep"]


class _CrossEntropyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, inv_count: float):
        eng = _engine.engine_for(logits)
        logits, labels = logits.contiguous(), labels.contiguous()
        loss, _nll, correct = eng.ce_fwd(logits, labels, inv_count)
        ctx.eng, ctx.inv_count = eng, inv_count
        ctx.save_for_backward(logits, labels)
        ctx.mark_non_differentiable(correct)
        return loss, correct

    @staticmethod
    def backward(ctx, gloss, _gcorrect):
        logits, labels = ctx.saved_tensors
        return ctx.eng.ce_bwd(logits, labels, ctx.inv_count, gloss.contiguous().float()), None, None


class CrossEntropyLoss(nn.CrossEntropyLoss):
    """Mean cross entropy over graphs on the device kernel.  ``forward(logits, labels)`` matches
    ``nn.CrossEntropyLoss()`` (default arguments); ``denominator`` overrides the divisor with the
    global batch size under data parallelism.  ``last_correct`` holds the argmax-accuracy count of
    the most recent call (device int64 scalar) - it falls out of the same pass."""

    def __init__(self):
        super().__init__()
        self.last_correct = None

    def forward(self, logits: torch.Tensor, labels: torch.Tensor, denominator: int | None = None) -> torch.Tensor:
        if logits.dim() != 2:
            raise ValueError("expected logits of shape [num_graphs, num_classes]")
        if labels is None or labels.dim() != 1 or labels.shape[0] != logits.shape[0]:
            # a partially labelled batch stacks fewer labels than subjects (reference graph.py:155-156,165); the
            # reference's nn.CrossEntropyLoss raises on that shape mismatch - so do we, instead of reading past the buffer
            got = "None" if labels is None else tuple(labels.shape)
            raise ValueError(f"expected one label per graph: logits {tuple(logits.shape)} but labels {got}")
        count = logits.shape[0] if denominator is None else denominator
        loss, correct = _CrossEntropyFn.apply(logits.float(), labels.to(torch.int64), 1.0 / max(count, 1))
        self.last_correct = correct
        return loss


class FlatParams:
    """One flat fp32 buffer for all trainable parameters of a classifier and one for their gradients.

    Every ``nn.Parameter`` keeps its identity (an optimizer built earlier stays valid) but its storage becomes a slice of
    ``self.param``; backward kernels write each gradient straight into the matching slice of ``self.grad`` and hand
    autograd a fresh view of it, so ``param.grad`` aliases the flat buffer without a copy.  The gradient all-reduce of a
    data-parallel step is then ONE collective over ``self.grad`` and the fused Adam update ONE kernel.
    Layout: per layer W, b, then the BatchNorm pair as [beta; gamma] (that [2, C] block is exactly the block of backward
    sums the kernels produce: sum dy, sum dy*xhat), then the head."""

    def __init__(self, model: nn.Module):
        order = []
        for conv, bn in zip(model.convs, model.batch_norms):
            order += [*conv.tensors(), bn.bias, bn.weight]
        fc0, fc1 = model.classifier[0], model.classifier[3]
        order += [fc0.weight, fc0.bias, fc1.weight, fc1.bias]
        seen = {id(q) for q in order}
        order += [q for q in model.parameters() if id(q) not in seen]
        total = sum(q.numel() for q in order)
        dev = order[0].device
        self.param = torch.empty(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.slots, off = {}, 0
        for q in order:
            n = q.numel()
            self.param[off:off + n].copy_(q.detach().reshape(-1))
            q.data = self.param[off:off + n].view(q.shape)
            self.slots[id(q)] = (off, n, tuple(q.shape))
            off += n
        self.params = order

    def view(self, q: torch.Tensor) -> torch.Tensor:
        """A NEW view of q's gradient slice (autograd adopts a tensor as ``.grad`` only if nobody else holds it)."""
        off, n, shape = self.slots[id(q)]
        return self.grad[off:off + n].view(shape)

    def bn_block(self, beta: torch.Tensor, gamma: torch.Tensor) -> torch.Tensor:
        off, n, _ = self.slots[id(beta)]
        assert self.slots[id(gamma)][0] == off + n
        return self.grad[off:off + 2 * n].view(2, n)


class GraphedStep:
    """A training or evaluation step of fixed shape captured as ONE CUDA graph: subject gather + collate, forward, loss,
    backward, gradient hand-over and the fused Adam update replay with a single launch (SURVEY 7 hard part f: at batch 16
    the ~35 kernels and their host-side launches ARE the step).  ``__call__(ids)`` copies the subject indices into the
    graph's input buffer and replays; it returns the loss (and the correct count for evaluation) as device scalars that the
    next replay overwrites."""

    def __init__(self, trainer: "Trainer", store, batch_size: int, kind: str, train: bool):
        import numpy as np
        self.trainer, self.train = trainer, train
        dev = trainer.device
        n0, e0 = int(store.node_ptr_host[1] - store.node_ptr_host[0]), int(store.edge_ptr_host[1] - store.edge_ptr_host[0])
        if not (np.all(np.diff(store.node_ptr_host) == n0) and np.all(np.diff(store.edge_ptr_host) == e0)):
            raise ValueError("a graphed step needs subjects of one size (the shapes of a captured launch are fixed)")
        if not bool(store.has_label.all()):
            raise ValueError("a graphed step needs every subject labelled")
        self.batch_size = batch_size
        self.ids_host = torch.zeros(batch_size, dtype=torch.int64).pin_memory()
        self.ids_dev = torch.zeros(batch_size, dtype=torch.int64, device=dev)
        ids0 = np.arange(batch_size, dtype=np.int64) % len(store)

        def body():
            batch = store.collate(ids0, ids_device=self.ids_dev, prepare_for=kind, backward=train)
            if train:
                return trainer.train_step(batch), None
            return trainer.eval_step(batch)

        model = trainer.model
        # the peer-memory SyncBN exchange takes a sequence number as a launch argument: a replayed graph would repeat it.
        # Captured data-parallel steps keep their BatchNorm exchanges on NCCL.
        model.peer_collectives = False
        model.train(train)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):          # warm-up outside the capture: lazy initialisation, allocator pools
            for _ in range(3):
                body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, self.correct = body()

    def __call__(self, ids):
        import numpy as np
        ids = np.asarray(ids, dtype=np.int64).reshape(-1)
        if ids.size != self.batch_size:
            raise ValueError(f"this step was captured for {self.batch_size} subjects, got {ids.size}")
        self.ids_host.copy_(torch.from_numpy(ids))
        self.ids_dev.copy_(self.ids_host, non_blocking=True)
        self.trainer.model.train(self.train)
        self.graph.replay()
        return (self.loss if self.train else (self.loss, self.correct))


def _dist_world() -> int:
    import torch.distributed as dist
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class Trainer:
    """Training loop for ConnectomeBatch classifiers: ``train_epoch``, ``evaluate``, ``fit``."""

    def __init__(self, model: nn.Module, optimizer: torch.optim.Optimizer, device: str = "cpu"):
        dev = torch.device(device)
        if dev.type != "cuda":
            dev = _engine.default_device()   # raises if no GPU: the B200 build never computes on the host
        self.model = model.to(dev)           # in place; the optimizer's Parameter objects stay valid
        self.optimizer = optimizer
        self.device = dev
        self.loss_fn = CrossEntropyLoss()
        self.distributed = True              # False: no collectives even when torch.distributed is initialised
        self.flat = None                     # FlatParams after enable_fused_step()
        self._adam = None                    # (exp_avg, exp_avg_sq, state, hyper-parameters) of the fused update

    # -- opt-in: flat buffers, one-kernel Adam, device-side step state -------------------------
    def enable_fused_step(self, seed: int | None = None) -> "Trainer":
        """Flat parameter / gradient buffers (``FlatParams``), and - when the optimizer is a plain ``torch.optim.Adam``
        with one parameter group - its update as ONE kernel over the flat buffers (``cgnn_adam_step``: the arithmetic of
        torch's Adam op for op) with the step counter and the dropout salt in a device state block, which is what lets
        ``capture`` replay whole steps as CUDA graphs.  Any other optimizer keeps working on the flat views."""
        if self.flat is not None:
            return self
        self.flat = FlatParams(self.model)
        self.model._flat = self.flat
        opt = self.optimizer
        plain = (type(opt) is torch.optim.Adam and len(opt.param_groups) == 1 and not opt.param_groups[0].get("amsgrad", False)
                 and not opt.param_groups[0].get("maximize", False) and len(opt.state) == 0 and
                 {id(q) for q in opt.param_groups[0]["params"]} == {id(q) for q in self.flat.params})
        eng = _engine.engine_for(self.flat.param)
        state = torch.zeros(4, dtype=torch.int64, device=self.device)      # cgnn_step_tick's block: step, seed, salt, spare
        state[1] = int(torch.initial_seed() if seed is None else seed) & 0x7FFFFFFFFFFFFFFF
        self._state = state
        if plain:
            g = opt.param_groups[0]
            self._adam = dict(m=torch.zeros_like(self.flat.param), v=torch.zeros_like(self.flat.param), lr=g["lr"],
                              b1=g["betas"][0], b2=g["betas"][1], eps=g["eps"], wd=g["weight_decay"], eng=eng)
        return self

    def _fused_optimizer_step(self) -> None:
        a = self._adam
        a["eng"].adam_step(self.flat.param, self.flat.grad, a["m"], a["v"], a["lr"], a["b1"], a["b2"], a["eps"], a["wd"], self._state)

    def capture(self, store, batch_size: int, prepare_for: str, train: bool = True) -> GraphedStep:
        """Capture ``train_step`` (or ``eval_step``) on batches of ``batch_size`` subjects of ``store`` as one CUDA graph.
        Training needs ``enable_fused_step()`` with a plain Adam (the update has to live on the device); the dropout masks
        of the replays come from the device salt, refreshed inside the graph."""
        if train:
            self.enable_fused_step()
            if self._adam is None:
                raise ValueError("a graphed training step needs a plain torch.optim.Adam (single group) to fuse")
            salt = self._state[2:3].view(torch.int32)         # the two 32-bit salt words cgnn_step_tick refreshes
            self.model._salt, self.model._graph_seed = salt, int(self._state[1].item())
        return GraphedStep(self, store, batch_size, prepare_for, train)

    def _world(self) -> int:
        return _dist_world() if self.distributed else 1

    # -- data-parallel plumbing ------------------------------------------------------------
    def _sync_gradients(self) -> None:
        """One flat all-reduce(sum) of every parameter gradient (loss was pre-scaled by 1/B_global)."""
        if self._world() == 1:
            return
        import torch.distributed as dist
        if self.flat is not None:          # the gradients already live in one buffer: one collective, nothing to copy
            dist.all_reduce(self.flat.grad)
            return
        params = [p for p in self.model.parameters() if p.requires_grad]
        for p in params:
            if p.grad is None:   # a rank with an empty slice still joins the collective
                p.grad = torch.zeros_like(p)
        flat = torch.cat([p.grad.reshape(-1) for p in params])
        dist.all_reduce(flat)
        offset = 0
        for p in params:
            n = p.numel()
            p.grad.copy_(flat[offset:offset + n].view_as(p))
            offset += n

    @staticmethod
    def _epoch_totals(values: list, weights: list):
        """Host reduction of per-batch device scalars: one device->host transfer per epoch."""
        if not values:
            return []
        return torch.stack([v.detach().reshape(()).double() for v in values]).cpu().tolist()

    # -- one batch (the loop bodies of reference train.py:45-53 and 61-68) ----------------------
    def train_step(self, batch) -> torch.Tensor:
        """zero_grad -> forward -> loss -> backward -> (gradient all-reduce) -> optimizer step.
        Returns this rank's share of the global mean loss as a device scalar (no host sync)."""
        batch = batch.to(self.device)
        global_b = batch.global_num_graphs if batch.global_num_graphs is not None else batch.num_graphs
        if self.flat is not None:
            _engine.engine_for(self.flat.param).step_tick(self._state)      # step counter + this step's dropout salt
            if batch.num_graphs == 0:
                self.flat.grad.zero_()      # an empty slice of a data-parallel batch contributes nothing
        self.optimizer.zero_grad()          # set_to_none: autograd then ADOPTS the gradient views instead of adding to them
        logits = self.model(batch)
        loss = self.loss_fn(logits, batch.labels, global_b)
        loss.backward()
        self._sync_gradients()
        if self._adam is not None:
            self._fused_optimizer_step()
        else:
            self.optimizer.step()
        return loss.detach()

    @torch.no_grad()
    def eval_step(self, batch):
        """forward -> loss and argmax-accuracy count, both device scalars (no host sync)."""
        batch = batch.to(self.device)
        global_b = batch.global_num_graphs if batch.global_num_graphs is not None else batch.num_graphs
        logits = self.model(batch)
        loss = self.loss_fn(logits, batch.labels, global_b)
        return loss, self.loss_fn.last_correct

    # -- reference API -----------------------------------------------------------------------
    def _lean(self, loader):
        """The loader's batches are consumed right here, by this trainer's model: unless the caller chose otherwise, a
        ``ConnectomeDataLoader`` then collates lean batches for that model family (only what its kernels read; the
        reference fields stay available on access).  Returns a context manager that undoes the setting."""
        import contextlib
        kind = getattr(self.model, "kind", None)

        @contextlib.contextmanager
        def scope():
            mine = kind in ("gcn", "sage") and hasattr(loader, "prepare_for") and getattr(loader, "prepare_for") is None
            if mine:
                loader.prepare_for = kind
            try:
                yield
            finally:
                if mine:
                    loader.prepare_for = None
        return scope()

    def train_epoch(self, loader) -> float:
        """One pass over ``loader`` with parameter updates; returns the mean loss (``train.py:41-54``)."""
        with self._lean(loader):
            return self._train_epoch(loader)

    def _train_epoch(self, loader) -> float:
        self.model.train()
        losses, sizes = [], []
        for batch in loader:
            losses.append(self.train_step(batch))
            sizes.append(batch.global_num_graphs if batch.global_num_graphs is not None else batch.num_graphs)
        if self._world() > 1 and losses:
            import torch.distributed as dist
            stacked = torch.stack(losses)
            dist.all_reduce(stacked)          # per-rank partial sums of nll / B_global -> global mean
            losses = list(stacked.unbind(0))
        total_loss, total_samples = 0.0, 0
        for value, b in zip(self._epoch_totals(losses, sizes), sizes):
            total_loss += float(value) * b
            total_samples += b
        return total_loss / max(total_samples, 1)

    @torch.no_grad()
    def evaluate(self, loader) -> dict:
        """Accuracy and mean loss over ``loader`` (``train.py:56-74``)."""
        with self._lean(loader):
            return self._evaluate(loader)

    def _evaluate(self, loader) -> dict:
        self.model.eval()
        losses, corrects, sizes = [], [], []
        for batch in loader:
            loss, correct_count = self.eval_step(batch)
            losses.append(loss)
            corrects.append(correct_count)
            sizes.append(batch.global_num_graphs if batch.global_num_graphs is not None else batch.num_graphs)
        if self._world() > 1 and losses:
            import torch.distributed as dist
            stacked = torch.stack([torch.stack(losses).double(), torch.stack(corrects).double()])
            dist.all_reduce(stacked)
            losses, corrects = list(stacked[0].unbind(0)), list(stacked[1].unbind(0))
        total_loss, correct, total = 0.0, 0, 0
        loss_vals = self._epoch_totals(losses, sizes)
        corr_vals = self._epoch_totals(corrects, sizes)
        for value, c, b in zip(loss_vals, corr_vals, sizes):
            total_loss += float(value) * b
            correct += int(round(c))
            total += b
        return {
            "accuracy": correct / max(total, 1),
            "loss": total_loss / max(total, 1),
            "correct": correct,
            "total": total,
        }

    def fit(self, train_loader, val_loader, num_epochs: int = 50, patience: int = 10, verbose: bool = True) -> dict:
        """Train with early stopping on the validation loss and restore the best weights
        (``train.py:76-127``: strict ``<`` improvement test, stop once ``patience`` epochs have
        passed since the best one, best state restored even without an early stop)."""
        history = {"train_loss": [], "val_loss": [], "val_acc": []}
        best = {"loss": float("inf"), "epoch": 0, "state": None}
        for epoch in range(1, num_epochs + 1):
            train_loss = self.train_epoch(train_loader)
            val = self.evaluate(val_loader)
            history["train_loss"].append(train_loss)
            history["val_loss"].append(val["loss"])
            history["val_acc"].append(val["accuracy"])
            if verbose:
                print(f"Epoch {epoch:3d} | train_loss={train_loss:.4f} | val_loss={val['loss']:.4f} | "
                      f"val_acc={val['accuracy']:.3f}")
            if val["loss"] < best["loss"]:
                best = {"loss": val["loss"], "epoch": epoch,
                        "state": {k: v.clone() for k, v in self.model.state_dict().items()}}
            if epoch - best["epoch"] >= patience:
                if verbose:
                    print(f"Early stop at epoch {epoch} (best={best['epoch']})")
                break
        if best["state"] is not None:
            self.model.load_state_dict(best["state"])
        return history
