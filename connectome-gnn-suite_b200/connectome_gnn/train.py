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

__all__ = ["Trainer", "CrossEntropyLoss"]


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

    def _world(self) -> int:
        return _dist_world() if self.distributed else 1

    # -- data-parallel plumbing ------------------------------------------------------------
    def _sync_gradients(self) -> None:
        """One flat all-reduce(sum) of every parameter gradient (loss was pre-scaled by 1/B_global)."""
        if self._world() == 1:
            return
        import torch.distributed as dist
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
        self.optimizer.zero_grad()
        logits = self.model(batch)
        loss = self.loss_fn(logits, batch.labels, global_b)
        loss.backward()
        self._sync_gradients()
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
    def train_epoch(self, loader) -> float:
        """One pass over ``loader`` with parameter updates; returns the mean loss (``train.py:41-54``)."""
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
