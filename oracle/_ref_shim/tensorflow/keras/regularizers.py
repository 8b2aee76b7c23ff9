"""``tensorflow.keras.regularizers`` of the shim.  TEST INFRASTRUCTURE ONLY."""


class L1L2:
    """``l2(l)``: penalty ``l * sum(w^2)`` added to the training loss (IL:217 puts
    ``l2(emb_reg)`` on every cross-embedding table, which makes the table gradient dense)."""

    def __init__(self, l1=0.0, l2=0.0):
        self.l1, self.l2 = float(l1 or 0.0), float(l2 or 0.0)

    def __call__(self, w):
        out = 0.0
        if self.l1:
            out = out + self.l1 * w.abs().sum()
        if self.l2:
            out = out + self.l2 * (w * w).sum()
        return out


def l2(l=0.01):
    return L1L2(l2=l)


def l1(l=0.01):
    return L1L2(l1=l)


def l1_l2(l1=0.01, l2=0.01):
    return L1L2(l1=l1, l2=l2)
