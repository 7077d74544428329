"""``create_model`` – mirror of models/generate_model.py:53-229 for the models on the
Log-ODE hot path (``log_ncde`` :86-109 and ``*_linear_ncde`` :110-129) and the ``nrde`` comparison model on the
same input (:151-172).  Same
signature, defaults and error behaviour; ``solver`` / ``stepsize_controller`` are
accepted for signature compatibility – the kernels implement Heun with a constant
step (the reference defaults, run_experiment.py:196-197)."""
from __future__ import annotations


def create_model(model_name, data_dim, logsig_dim, logsig_depth, intervals, label_dim, hidden_dim,
                 num_blocks=None, block_size=None, vf_depth=None, vf_width=None, classification=True,
                 output_step=1, ssm_dim=None, ssm_blocks=None, solver="Heun", stepsize_controller="ConstantStepSize",
                 dt0=1, max_steps=16**4, scale=1.0, lambd=0.0, w_init_std=0.25, parallel_steps=1,
                 walsh_hadamard=False, diagonal_dense=False, sparsity=1.0, rank=0, *, key, device="cuda"):
    if model_name == "log_ncde":
        if vf_width is None or vf_depth is None:
            raise ValueError("Must specify vf_width and vf_depth for a Log-NCDE.")
        from .LogNeuralCDEs import LogNeuralCDE

        return (LogNeuralCDE(vf_width, vf_depth, hidden_dim, data_dim, logsig_depth, label_dim, classification,
                             output_step, intervals, solver, stepsize_controller, dt0, max_steps, scale, lambd,
                             key=key, device=device), None)
    if model_name.endswith("_linear_ncde"):
        from .LinearNeuralCDEs import LogLinearCDE

        return (LogLinearCDE(data_dim=data_dim, hidden_dim=hidden_dim, label_dim=label_dim, block_size=block_size,
                             logsig_depth=logsig_depth, lambd=lambd, w_init_std=w_init_std,
                             classification=classification, parallel_steps=parallel_steps,
                             walsh_hadamard=walsh_hadamard, diagonal_dense=diagonal_dense, sparsity=sparsity, rank=rank,
                             key=key, device=device),
                None)
    if model_name == "nrde":
        if vf_width is None or vf_depth is None:
            raise ValueError("Must specify vf_width and vf_depth for a NRDE.")
        from .NeuralCDEs import NeuralRDE

        return (NeuralRDE(vf_width, vf_depth, hidden_dim, data_dim, logsig_dim, label_dim, classification, output_step,
                          intervals, solver, stepsize_controller, dt0, max_steps, scale, key=key, device=device), None)
    if model_name in ("ncde", "lru", "S5", "rnn_linear", "rnn_gru", "rnn_lstm", "rnn_mlp"):
        raise NotImplementedError(f"{model_name} is a comparison baseline outside the Log-ODE hot path (SURVEY.md §2)")
    raise ValueError(f"Unknown model name: {model_name}")
