"""Drop-in replacement of the reference ``gnn_data.py`` (synthetic sparse-recovery data)."""
import torch
from torch.utils.data import Dataset, DataLoader


class GNN_Data(Dataset):
    """(b [N,P,m,1], y [N,n,1]) pairs (reference gnn_data.py:18-27)."""

    def __init__(self, b, y):
        self.b, self.y = b, y

    def __len__(self):
        return self.b.shape[0]

    def __getitem__(self, item):
        return self.b[item], self.y[item]


def make_problem(A, data_len):
    """Labels y = 2*N(0,1)*Bernoulli(0.25) and noiseless observations b_p = A_p y (reference :6-14:
    the sigma-scaled noise drawn at :12 is overwritten at :13-14, so it only advances the RNG).

    With ``A`` on a CUDA device the P per-agent products of the reference's loop (:13-14) are ONE launch of the library's
    contraction (every agent reads the same y: a zero agent stride, no [N,P,n] copy); on the CPU the loop stays as it is."""
    device = A.device
    _, P, m, n = A.shape
    y = 2 * torch.randn(data_len, n, 1, device=device) * (torch.rand(data_len, n, 1, device=device) <= 0.25)
    b = torch.randn(data_len, P, m, 1, device=device)          # keeps the reference's RNG stream position
    if A.is_cuda:
        from dadmm_b200 import functional as DF
        DF.observe(A[0], y.squeeze(-1), out=b.squeeze(-1))
        return b, y
    for p in range(P):
        b[:, p] = torch.matmul(A[0, p], y)
    return b, y


def set_Data(A, data_len, args):
    b, y = make_problem(A, data_len)
    return DataLoader(GNN_Data(b, y), batch_size=args.batch_size, shuffle=True, drop_last=True)
