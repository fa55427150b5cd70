"""TEST INFRASTRUCTURE (never imported by the product): CPU restatement of the data-parallel optimizer step of
bert_multimodal_transformer_b200/csrc/peer_optim.cu as the trainer schedules it (trainer._peer_step), in plain torch.

What it replaces in the reference: DistributedDataParallel's gradient all-reduce around `loss.backward()` plus
`optimizer.step()` with transformers-3.0.2 AdamW (multimodal_driver.py:378-386).  The restatement keeps the structure of
the kernels -- ownership by bucket piece, the sum over ranks in rank order, the early pass over the word-embedding rows no
rank touched (gradient exactly zero), the late pass over touched rows and everything else, parameters copied to every
rank -- so that tests can check the SCHEDULE against the textbook form (mean gradient, then AdamW everywhere)."""
import math

import torch


def owned_piece(begin, end, world, rank):
    """parallel.owned_piece restated: 256-aligned 1 / world of [begin, end)."""
    chunk = -(-(end - begin) // (world * 256)) * 256
    b = min(end, begin + rank * chunk)
    return b, min(end, b + chunk)


def adamw_(p, g, m, v, decay, lr, b1, b2, eps, wd, step):
    """transformers 3.0.2 AdamW on flat slices, in place (head_optim.cu::adamw_kernel / adamw_peer_element)."""
    m.mul_(b1).add_(g, alpha=1.0 - b1)
    v.mul_(b2).addcmul_(g, g, value=1.0 - b2)
    step_size = lr * math.sqrt(1.0 - b2 ** step) / (1.0 - b1 ** step)
    p.addcdiv_(m, v.sqrt().add_(eps), value=-step_size)
    p.sub_(p * (decay * (lr * wd)))


def peer_step(params, grads, moments, decay, table, token_ids, lr, b1, b2, eps, wd, step):
    """One optimizer step over `world = len(params)` replicas.

    params[r], grads[r]: flat fp32 tensors of rank r (replicas of the same parameters, different gradients)
    moments[r] = (m, v): rank r's Adam moments (only the pieces rank r owns are ever touched)
    decay: flat 0/1 tensor (weight decay applies); table = (begin, rows, row_len) of the word-embedding matrix
    token_ids[r]: the ids of rank r's batch.  Gradients are cleared at the end, as the clean-up launch does."""
    world, n = len(params), params[0].numel()
    t0, rows, rl = table
    any_row = torch.zeros(rows, dtype=torch.bool)
    for ids in token_ids:                                   # peer_mark_rows + peer_combine_marks
        any_row[ids.reshape(-1)] = True
    touched = torch.zeros(n, dtype=torch.bool)
    touched[t0:t0 + rows * rl] = any_row.repeat_interleave(rl)
    in_table = torch.zeros(n, dtype=torch.bool)
    in_table[t0:t0 + rows * rl] = True
    lo0 = t0 + rows * rl                                    # bucket 0 = the table (+ nothing else here), bucket 1 = the rest
    lo0 = (lo0 + 255) // 256 * 256
    buckets = [(0, lo0), (lo0, n)]

    def run(rank, sel, with_grads):
        m, v = moments[rank]
        idx = sel.nonzero().reshape(-1)
        if idx.numel() == 0:
            return
        g = torch.zeros(idx.numel())
        if with_grads:
            for q in range(world):                          # fixed rank order
                g = g + grads[q][idx]
            g = g * (1.0 / world)
        p, mm, vv = params[rank][idx].clone(), m[idx].clone(), v[idx].clone()
        adamw_(p, g, mm, vv, decay[idx], lr, b1, b2, eps, wd, step)
        m[idx], v[idx] = mm, vv
        for q in range(world):                              # all-gather: every rank receives the owner's result
            params[q][idx] = p

    # early pass (before backward in the real schedule): untouched table rows, gradient exactly zero
    for r in range(world):
        b, e = owned_piece(*buckets[0], world, r)
        own = torch.zeros(n, dtype=torch.bool)
        own[b:e] = True
        run(r, own & in_table & ~touched, with_grads=False)
    # late pass: touched rows, the non-table part of bucket 0, bucket 1
    for r in range(world):
        for bk in buckets:
            b, e = owned_piece(*bk, world, r)
            own = torch.zeros(n, dtype=torch.bool)
            own[b:e] = True
            run(r, own & ~(in_table & ~touched), with_grads=True)
    for g in grads:
        g.zero_()
