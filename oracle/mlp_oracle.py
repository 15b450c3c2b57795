"""CPU oracle for the MLP back-constraint of a hidden layer (second half).
TEST INFRASTRUCTURE ONLY (see oracle/psi_oracle.py).

Restates with numpy and the reference's loops:
  * the network of autoreg/mlp.py:13-93, 115-160: layers ``out = act(in W^T + b)`` (:66), tanh on
    the hidden layers, linear on the last (``positive_obs=False``, :127), default widths
    [Q, 2 Q, Q + X_dim / 2, X_dim] (autoreg/layers.py:441); ``update_gradient`` (:134-142) = gradient
    of sum(external_grad * output) with respect to the input, accumulating the W / b gradients;
  * ``_encoder_freerun``  autoreg/layers.py:623-666: the first X_win latent means are free
    parameters (init_Xs), every later mean is the network's output for the window before it
    (and the aligned control window);
  * ``_encoder_update_gradient``  :668-715: back-propagation through that recurrence, newest step
    first, adding each step's input gradient onto the latent / control means it was built from.
The reference evaluates the network with theano, which is absent here: parity is pinned by finite
differences of a scalar functional of the free-run output (tests/test_mlp_oracle.py).
"""
import numpy as np


def default_units(Q, X_dim):
    return [Q, Q * 2, Q + X_dim // 2, X_dim]                       # layers.py:441 (Python-2 integer division)


def mlp_forward(params, x):
    """params = [(W [down, up], b [down]), ...]; returns (output, activations incl. the input)."""
    acts = [np.asarray(x, dtype=np.float64)]
    for l, (W, b) in enumerate(params):
        s = acts[-1] @ W.T + b                                      # mlp.py:66
        acts.append(np.tanh(s) if l < len(params) - 1 else s)       # :67-72, :127
    return acts[-1], acts


def mlp_backward(params, acts, dL):
    """Gradient of sum(dL * output): returns (d/d input, [(dW, db), ...]) (mlp.py:134-142)."""
    grads = [None] * len(params)
    delta = np.asarray(dL, dtype=np.float64)
    for l in range(len(params) - 1, -1, -1):
        W, _ = params[l]
        if l < len(params) - 1:
            delta = delta * (1.0 - acts[l + 1] ** 2)
        grads[l] = (np.outer(delta, acts[l]), delta.copy())
        delta = delta @ W
    return delta, grads


def freerun(params, init_means, ctl_means, n_steps, X_win, U_win):
    """layers.py:623-666 for every sequence: init_means[s] [X_win, X_dim], ctl_means[s] [T_u, U_dim]
    or None, n_steps[s] = N.  Returns the latent means [X_win + N, X_dim] per sequence."""
    out = []
    for s, init in enumerate(init_means):
        N, X_dim = n_steps[s], init.shape[1]
        X = np.zeros((X_win + N, X_dim))
        X[:X_win] = init
        U = ctl_means[s] if ctl_means is not None else None
        U_off = U.shape[0] - N - U_win + 1 if U is not None else 0
        for n in range(N):
            x_in = X[n:n + X_win].ravel()
            if U is not None:
                x_in = np.concatenate([x_in, U[U_off + n:U_off + n + U_win].ravel()])
            X[X_win + n] = mlp_forward(params, x_in)[0]
        out.append(X)
    return out


def freerun_backward(params, lat_means, ctl_means, lat_grads, X_win, U_win):
    """layers.py:668-715.  lat_grads[s] [X_win + N, X_dim] are dL/d mean of every step (from the
    bound); they are updated in place, and on return their first X_win rows are the gradients of
    the initial means.  Returns ([(dW, db), ...] summed over sequences, control-mean gradients)."""
    total = [(np.zeros_like(W), np.zeros_like(b)) for W, b in params]
    ctl_grads = [np.zeros_like(u) for u in ctl_means] if ctl_means is not None else None
    for s, X in enumerate(lat_means):
        N, X_dim = X.shape[0] - X_win, X.shape[1]
        U = ctl_means[s] if ctl_means is not None else None
        U_off = U.shape[0] - N - U_win + 1 if U is not None else 0
        g = lat_grads[s]
        for n in range(N - 1, -1, -1):
            x_in = X[n:n + X_win].ravel()
            if U is not None:
                x_in = np.concatenate([x_in, U[U_off + n:U_off + n + U_win].ravel()])
            _, acts = mlp_forward(params, x_in)
            dX, grads = mlp_backward(params, acts, g[X_win + n])
            for l, (dW, db) in enumerate(grads):
                total[l][0][...] += dW
                total[l][1][...] += db
            g[n:n + X_win] += dX[:X_win * X_dim].reshape(-1, X_dim)
            if U is not None:
                ctl_grads[s][U_off + n:U_off + n + U_win] += dX[X_win * X_dim:].reshape(-1, U.shape[1])
    return total, ctl_grads
