"""ORACLE -- TEST INFRASTRUCTURE ONLY (tests/ may import this; the product never does).

numpy restatement of the training step of the reference (training.rs:64-67,137-200,277-292,424-440) in float64:
  compute_gradients   loss = mean_b(-sum_i pi*log(p + 1e-5)) + 0.5 * mean_b((v - z)^2)          training.rs:277-292
  gradient clipping   GradientClippingConfig::Value(1.0): every element clamped to [-1, 1]        training.rs:65
  AdamW               burn 0.18 AdamW (adaptive moments with bias correction, decoupled decay), defaults beta1 0.9,
                      beta2 0.999, epsilon 1e-5, weight decay 1e-4 (parameters.rs:26)              training.rs:64-67
  get_cyclical_lr     triangular 1e-3 <-> 1e-2 over 20 iterations, x0.1 every 1000                 training.rs:424-440
Parity status: burn is not vendored in /root/reference, so the optimizer's arithmetic is restated from burn's published
algorithm [recalled]: theta <- theta - lr*wd*theta - lr * m_hat / (sqrt(v_hat) + eps); "parity unpinned" by the reference.
"""
import numpy as np

BETA1, BETA2, EPS, WEIGHT_DECAY = 0.9, 0.999, 1e-5, 1e-4
VALUE_LOSS_WEIGHT = 0.5


def loss(pred_policy, target_policy, pred_value, target_value):
    p, pi = np.asarray(pred_policy, np.float64), np.asarray(target_policy, np.float64)
    v, z = np.asarray(pred_value, np.float64), np.asarray(target_value, np.float64)
    policy_loss = float(np.mean(-(pi * np.log(p + 1e-5)).sum(1)))
    value_loss = float(np.mean((v - z) ** 2))
    return policy_loss, value_loss, policy_loss + VALUE_LOSS_WEIGHT * value_loss


def loss_gradients(pred_policy, target_policy, pred_value, target_value):
    """d loss / d pred_policy and d loss / d pred_value (what autograd starts from)."""
    p, pi = np.asarray(pred_policy, np.float64), np.asarray(target_policy, np.float64)
    v, z = np.asarray(pred_value, np.float64), np.asarray(target_value, np.float64)
    n = p.shape[0]
    return -(pi / (p + 1e-5)) / n, VALUE_LOSS_WEIGHT * 2.0 * (v - z) / n


def cyclical_lr(iteration):
    decay = 10.0 ** (-(iteration // 1000))
    base, top = 1e-3 * decay, 1e-2 * decay
    cur = iteration % 20
    if cur <= 10:
        return base + (cur / 10) * (top - base)
    return top - ((cur - 10) / 10) * (top - base)


class AdamW:
    def __init__(self, shapes):
        self.m = [np.zeros(s, np.float64) for s in shapes]
        self.v = [np.zeros(s, np.float64) for s in shapes]
        self.t = 0

    def step(self, params, grads, lr):
        """params, grads: lists of float64 arrays; returns the updated parameters (gradients clipped by value to +-1)."""
        self.t += 1
        out = []
        for i, (w, g) in enumerate(zip(params, grads)):
            g = np.clip(np.asarray(g, np.float64), -1.0, 1.0)
            self.m[i] = BETA1 * self.m[i] + (1 - BETA1) * g
            self.v[i] = BETA2 * self.v[i] + (1 - BETA2) * g * g
            m_hat = self.m[i] / (1 - BETA1 ** self.t)
            v_hat = self.v[i] / (1 - BETA2 ** self.t)
            w = np.asarray(w, np.float64)
            out.append(w - lr * WEIGHT_DECAY * w - lr * m_hat / (np.sqrt(v_hat) + EPS))
        return out
