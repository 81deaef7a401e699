"""Epoch driver of the reference scripts, restated once for all five of them and
fixed for current torch:  ``run()`` = Adam/AdamW + ``ReduceLROnPlateau(factor=0.1, patience=p)`` on
the validation loss + best-checkpoint saving + early stop after 4 epochs without a new best
(others/realformer.py:338-363, cmu-mosei/run.py:394-419, Ren-MME/run.py:370-401,
rencecps/run.py:198-223, robot_demo.py:498-523), and the contiguous k-fold split the scripts spell
out by hand (others/realformer.py:365-389).

Differences from the reference, all outside the hot path:
* ``ReduceLROnPlateau(verbose=True)`` raises ``TypeError`` on torch >= 2.7; the lr changes are
  reported through ``log`` instead.
* the model / batch / loss specifics live in two callables (``train_step``, ``valid_step``) so the
  same driver serves every script; with ``mmemo_b200.optim.AdamW(max_grad_norm=CLIP)`` the clip is
  part of ``optimizer.step()``.
* TensorBoard logging is optional (``writer`` may be any object with ``add_scalars``).

Host-side orchestration only: nothing here touches the GPU directly.
"""
from __future__ import annotations

import os
from typing import Callable, Iterable, List, Optional, Sequence, Tuple

import torch
from torch.optim.lr_scheduler import ReduceLROnPlateau


def kfold_splits(names: Sequence, k: int = 5) -> List[Tuple[list, list]]:
    """[(train_list, valid_list)] with validation fold i = ``names[int(n*i/k):int(n*(i+1)/k)]``
    — the slicing of others/realformer.py:366-389 (k = 5, 20 % folds; the last fold runs to the
    end of the list)."""
    n = len(names)
    names = list(names)
    out = []
    for i in range(k):
        lo = int(n * (i / k)) if i else 0
        hi = int(n * ((i + 1) / k)) if i + 1 < k else n
        out.append((names[:lo] + names[hi:], names[lo:hi]))
    return out


def train_epoch(model: torch.nn.Module, iterator: Iterable, optimizer: torch.optim.Optimizer,
                train_step: Callable, clip: Optional[float] = 1.0) -> float:
    """others/realformer.py:300-318: mean of the per-batch losses.  ``train_step(model, batch)``
    returns the scalar loss tensor; ``clip=None`` when the optimizer clips itself."""
    model.train()
    total, count = 0.0, 0
    for batch in iterator:
        count += 1
        optimizer.zero_grad()
        loss = train_step(model, batch)
        loss.backward()
        if clip is not None:
            torch.nn.utils.clip_grad_norm_(model.parameters(), clip)
        optimizer.step()
        total += loss.item()
    return total / max(count, 1)


def valid_epoch(model: torch.nn.Module, iterator: Iterable, valid_step: Callable) -> float:
    """others/realformer.py:320-336."""
    model.eval()
    total, count = 0.0, 0
    with torch.no_grad():
        for batch in iterator:
            count += 1
            total += float(valid_step(model, batch))
    return total / max(count, 1)


def checkpoint_name(name: str, valid_loss: float) -> str:
    """``model_1_1.31.pt``: first four characters of ``str(valid_loss)`` (others/realformer.py:359)."""
    return f"{name}_{str(valid_loss)[:4]}.pt"


def fit(model: torch.nn.Module, optimizer: torch.optim.Optimizer,
        make_train_iter: Callable[[], Iterable], make_valid_iter: Callable[[], Iterable],
        train_step: Callable, valid_step: Optional[Callable] = None, *, epochs: int, name: str,
        log_dir: Optional[str] = None, clip: Optional[float] = 1.0, sched_patience: int = 2,
        sched_factor: float = 0.1, stop_after: int = 4, writer=None,
        log: Callable[[str], None] = print) -> dict:
    """The reference's ``run()``.  Returns ``{"train": [...], "valid": [...], "best": path | None,
    "lrs": [...]}``.  A new iterator is built every epoch (the reference reshuffles in
    ``data_loader``).  ``sched_patience`` is 2 in realformer and 1 in the other scripts."""
    valid_step = valid_step or (lambda m, b: train_step(m, b))
    scheduler = ReduceLROnPlateau(optimizer, factor=sched_factor, patience=sched_patience)
    log_file = None
    if log_dir is not None:
        os.makedirs(log_dir, exist_ok=True)
        log_file = os.path.join(log_dir, name + ".txt")
        with open(log_file, "w") as fh:
            fh.write("epoch, train_loss, valid_loss\n")
    hist = {"train": [], "valid": [], "best": None, "lrs": []}
    stop = 0
    for epoch in range(epochs):
        log(f"Epoch: {epoch + 1}")
        tr = train_epoch(model, make_train_iter(), optimizer, train_step, clip)
        va = valid_epoch(model, make_valid_iter(), valid_step)
        if writer is not None:
            writer.add_scalars(name, {"train_loss": tr, "valid_loss": va}, epoch)
        before = [g["lr"] for g in optimizer.param_groups]
        scheduler.step(va)
        after = [g["lr"] for g in optimizer.param_groups]
        if after != before:
            log(f"reducing learning rate: {before} -> {after}")
        hist["train"].append(tr)
        hist["valid"].append(va)
        hist["lrs"].append(after[0])
        if log_file is not None:
            with open(log_file, "a") as fh:
                fh.write("\n{epoch},{tr: 2.2f},{va: 2.2f}\n".format(epoch=epoch + 1, tr=tr, va=va))
        if va == min(hist["valid"]):
            stop = 0
            if log_dir is not None:
                hist["best"] = os.path.join(log_dir, checkpoint_name(name, va))
                torch.save(model.state_dict(), hist["best"])
        else:
            stop += 1
            if stop >= stop_after:
                break
    return hist
