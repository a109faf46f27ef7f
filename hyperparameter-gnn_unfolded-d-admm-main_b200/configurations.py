"""Drop-in replacement of the reference ``configurations.py``: the same command-line flags, types and
defaults (reference configurations.py:3-131), declared as a table."""
import argparse

# (flag, kwargs) in the reference's order
_FLAGS = [
    ("--m", dict(type=int, default=100, help="rows per agent")),
    ("--n", dict(type=int, default=500, help="number of unknowns")),
    ("--alpha_max", dict(type=float, default=0.1, help="upper bound of the step size alpha")),
    ("--tau_max", dict(type=float, default=0.99, help="upper bound of the l1 weight tau")),
    ("--rho_max", dict(type=float, default=0.99, help="upper bound of the consensus penalty rho")),
    ("--eta_max", dict(type=float, default=0.99, help="upper bound of the dual step eta")),
    ("--init_alpha_frac", dict(type=float, default=0.2, help="initial alpha / alpha_max")),
    ("--init_tau_frac", dict(type=float, default=0.15, help="initial tau / tau_max")),
    ("--init_rho_frac", dict(type=float, default=0.25, help="initial rho / rho_max")),
    ("--init_eta_frac", dict(type=float, default=0.1, help="initial eta / eta_max")),
    ("--max_penalty_threshold", dict(type=float, default=0.8, help="mean hyper-parameter level that triggers the penalty")),
    ("--penalty_reduction_factor", dict(type=float, default=0.95, help="shrink factor applied by the penalty")),
    ("--exp_name", dict(type=str, default="exp for 5 agents", help="experiment name")),
    ("--eval", dict(action="store_true", help="inference only")),
    ("--method", dict(type=str, default="u-dadmm", help="dadmm or u-dadmm")),
    ("--seq_num", dict(type=int, default=0, help="sequential-training segment index")),
    ("--data", dict(type=str, default="simulated", choices=["mnist", "simulated"], help="dataset")),
    ("--norm_mean", dict(type=float, default=0.5, help="normalisation mean")),
    ("--norm_std", dict(type=float, default=0.5, help="normalisation std")),
    ("--train_size", dict(type=int, default=200, help="training-set size")),
    ("--snr", dict(type=int, default=4, help="SNR of the simulated data")),
    ("--test_size", dict(type=int, default=32, help="test-set size")),
    ("--batch_size", dict(type=int, default=16, help="batch size")),
    ("--P", dict(type=int, default=5, help="number of agents")),
    ("--graph_prob", dict(type=float, default=0.5, help="Erdos-Renyi edge probability")),
    ("--graph_type", dict(type=str, default="erods_renyi", help="graph family")),
    ("--case", dict(type=str, default="dlasso", choices=["dlasso", "dlr"], help="problem")),
    ("--model", dict(type=str, default="same", choices=["diff", "same"], help="shared or per-agent hyper-parameters")),
    ("--rho", dict(type=float, default=0.2603, help="fixed rho (classic D-ADMM)")),
    ("--alpha", dict(type=float, default=0.3013, help="fixed alpha (classic D-ADMM)")),
    ("--eta", dict(type=float, default=0.0867, help="fixed eta (classic D-ADMM)")),
    ("--gamma", dict(type=float, default=1.1797e-07, help="fixed gamma (dlr)")),
    ("--beta", dict(type=float, default=1.2260e-03, help="fixed beta (dlr)")),
    ("--delta", dict(type=float, default=1.2665e-04, help="fixed delta (dlr)")),
    ("--tau", dict(type=float, default=0.1142, help="fixed tau (dlasso)")),
    ("--sequential", dict(type=bool, default=False, help="sequential training")),
    ("--max_iter_seg", dict(type=int, default=2, help="iterations per training segment")),
    ("--max_iter", dict(type=int, default=25, help="D-ADMM iterations")),
    ("--num_epochs", dict(type=int, default=10, help="epochs")),
    ("--optimizer", dict(type=str, default="adam", choices=["sgd", "adam"], help="optimizer")),
    ("--lr", dict(type=float, default=1e-04, help="learning rate")),
    ("--momentum", dict(type=float, default=0.5 * 1e-05, help="momentum")),
    ("--weight_decay", dict(type=float, default=0.0001, help="weight decay")),
    ("--lr_scheduler", dict(action="store_true", help="reduce lr on plateau")),
    ("--device", dict(type=str, default="cpu", choices=["cuda:0", "cuda:1", "cpu"], help="device")),
    ("--valid", dict(type=bool, default=True, help="run validation")),
    ("--seed", dict(type=float, default=42, help="manual seed")),
    ("--GHyp_hidden", dict(type=float, default=100, help="hypernetwork hidden width")),
    ("--DADMM_mode", dict(type=str, default="diff", choices=["same", "diff"], help="per-agent or shared hyper-parameters")),
    ("--hyp_mode", dict(type=str, default="unfolded", choices=["GHyp", "unfolded"], help="hyper-parameter source")),
    ("--GHN_iter_num", dict(type=int, default=15, help="unfolded D-ADMM iterations K")),
    ("--save_dir", dict(type=str, default="./results", help="output directory")),
]


def build_parser():
    parser = argparse.ArgumentParser()
    for flag, kw in _FLAGS:
        parser.add_argument(flag, **kw)
    return parser


def args_parser(argv=None):
    return build_parser().parse_args(argv)
