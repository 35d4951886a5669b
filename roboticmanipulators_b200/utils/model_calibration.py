"""ModelCalibrator: drop-in for utils/model_calibration.py of the reference.

Host side (pure Python, runs once): the irreducible-model builder `get_reduced_model`
(model_calibration.py:122-266) and, per iteration, the "compilation" of the reduced chain
at the current parameters into the POD block the kernels consume (constant token runs
pre-multiplied, per-parameter screw axis/origin in its joint frame).

Device side: the per-measurement loops of `estimate_parameters` (J^T J / J^T r),
`estimate_tool_base` (A^T A / A^T dp) and `evaluate` are single CUDA launches
(rk_calib_normal_eq_f64, rk_calib_tool_base_eq_f64, rk_calib_eval_f64) over measurements
resident on the GPU.  With torch.distributed initialised each rank passes its shard of the
measurements and the packed normal equations (190 doubles for the Fanuc model) are summed
with ONE all-reduce; the 18x18 solve is done redundantly on every rank, on the host, with
the reference's own np.linalg.inv(H).dot(g).
"""
import ctypes as C

import numpy as np
import sympy as sp
import torch
from scipy.optimize import least_squares

from .. import _lib
from .jacobians import JacobianCalculator
from .robo_math import TOKENS, Transformation as tf, chain_batch, elementary_np, tokens_from_sequence

_FULL_LINK = ("Tx", "Ty", "Tz", "Rx", "Ry", "Rz")


class ModelCalibrator(object):
    """Irreducible geometric model + least-squares estimation of T_base, T_tool and the
    geometric parameters from (joint configuration, measured points) pairs."""

    def __init__(self, sequence_string, joint_indices, variables,
                 link_lengths, offsets=None, directions=None, step=0.001):
        self._seq = sequence_string
        self._tokens = tokens_from_sequence(self._seq)
        self._ind = joint_indices
        self._jc = None
        self._parameters = None
        self._base_est = None
        self._T_base = None
        self._T_tools = None
        self._step = step
        self._no_model_msg = """
        Please, create the model first with ModelCalibrator.get_reduced_model()
        """
        self._base_iteration = 0
        self._params_iteration = 0
        self.verbose = True
        self._ws = {}

        self.set_link_lengths(link_lengths)
        self.set_joint_offsets(offsets)
        self.set_joint_directions(directions)

        if variables is None:
            self._variables, ti, ri = [], 0, 0
            for token in self._tokens:
                if token not in TOKENS:
                    raise ValueError("Unknown transformation")
                if token.startswith('T'):
                    self._variables.append(f"d_{ti}")
                    ti += 1
                else:
                    self._variables.append(f"q_{ri}")
                    ri += 1
        else:
            if len(variables) != len(self._tokens):
                raise ValueError("Transformation and var sizes do not match")
            self._variables = variables

        # joint-link groups: every joint token with the constant tokens that follow it
        # (model_calibration.py:87-93)
        starts = [i for i, x in enumerate(list(self._ind) + [1]) if x == 1]
        self._j_seq, self._j_var, self._j_ind = [], [], []
        for a, b in zip(starts, starts[1:]):
            self._j_seq.append(''.join(self._tokens[a:b]))
            self._j_var.append(self._variables[a:b])
            self._j_ind.append(self._ind[a:b])
        first_j = list(self._ind).index(1)
        last_j = len(self._ind) - 1 - list(self._ind)[::-1].index(1)
        self._to_base = self._tokens[:first_j]
        self._to_tool = self._tokens[last_j + 1:]
        self._to_base_vars = self._variables[:first_j]
        self._to_tool_vars = self._variables[last_j + 1:]

    def set_link_lengths(self, link_lengths):
        if link_lengths is None:
            self._link_lengths = np.ones(len(self._ind) - sum(self._ind))
        else:
            self._link_lengths = link_lengths

    def set_joint_offsets(self, offsets):
        self.qs_offsets = np.zeros(sum(self._ind)) if offsets is None else offsets

    def set_joint_directions(self, directions):
        self.qs_directions = np.ones(sum(self._ind)) if directions is None else directions

    # ----------------------------------------------------------------- model reduction
    def get_reduced_model(self):
        """Build the irreducible model (model_calibration.py:122-266).

        Every joint gets the full 6-parameter link after it minus what is redundant with
        the joint itself; then parameters that cannot be told apart from those of the
        neighbouring joint (consecutive parallel / orthogonal axes), the translations that
        slide into the tool, and the last link (absorbed by the tool) are removed.
        Returns (sequence string, variable names)."""
        groups = [tokens_from_sequence(s) for s in self._j_seq]
        jl = []                                    # per joint: [joint token, link tokens...]
        for g in groups:
            kind, ax = g[0][0], g[0][1]
            if kind == 'R':
                link = [t for t in _FULL_LINK if t not in (f"R{ax}", f"T{ax}")]
            else:
                link = [t for t in _FULL_LINK[3:] if t != f"R{ax}"]
            jl.append([f"{kind}{ax}"] + link)

        def drop(seq, token):
            return [t for t in seq if t != token]

        for j in range(1, len(jl)):
            cur, prev = jl[j], jl[j - 1]
            if cur[0][0] != prev[0][0]:
                continue
            kind, ax = cur[0][0], cur[0][1]
            if ax != prev[0][1]:                   # consecutive orthogonal axes
                jl[j - 1] = drop(prev, f"{kind}{ax}")
                continue
            k = 1                                  # consecutive parallel axes
            while j - k >= 0:
                cand = jl[j - k]
                if kind == 'R':
                    if cand[0][0] == 'T':
                        k += 1
                        continue
                    link_axis = groups[k][1][1] if len(groups[k]) > 1 else ''  # sic: group k
                    free = [a for a in "xyz" if a != cand[0][1] and a != link_axis]
                    jl[j - k] = drop(cand, f"T{free[0]}")
                    break
                if cand[0][0] == 'T' or cand[0][1] == ax:
                    k += 1
                    continue
                jl[j - k] = drop(cand, f"T{ax}")
                break

        barriers = set()                           # translations that slide into the tool
        for j in range(len(jl) - 1, 0, -1):
            barriers.add(jl[j][0][1])
            if len(barriers) > 1:
                break
            if groups[j - 1][0][0] != 'T':
                jl[j - 1] = drop(jl[j - 1], f"T{jl[j][0][1]}")
        jl[-1] = jl[-1][:1]                        # last link goes to the tool

        tokens, variables, nominal_and_qs, joint_indices = [], [], [], []
        for i, (seq, orig, names) in enumerate(zip(jl, groups, self._j_var), start=1):
            joint, link = seq[0], seq[1:]
            joint_indices.append(1)
            nominal_and_qs.append(names[0])
            tokens.append(joint)
            variables.append(names[0])
            if 1 < i < len(jl):                    # inner joints get an offset parameter
                tokens.append(joint)
                variables.append(f"d_{names[0]}")
                nominal_and_qs.append(0)
                joint_indices.append(0)
            orig_link, link_names = orig[1:], names[1:]
            for tok in link:
                tokens.append(tok)
                joint_indices.append(0)
                variables.append(f"{'phi' if tok[0] == 'R' else 'p'}_{tok[1]}_{i}")
                nominal_and_qs.append(link_names[orig_link.index(tok)] if tok in orig_link else 0)

        self._r_seq = ''.join(tokens)
        self._r_var = variables
        self._r_par = [v for v, i in zip(variables, joint_indices) if i != 1]
        self._parameters = np.zeros(len(self._r_par))
        self._r_nom_par = [v for v, i in zip(nominal_and_qs, joint_indices) if i != 1]
        self._r_nom_and_q = nominal_and_qs
        self._r_j_ind = joint_indices
        self._r_par_ind = 1 - np.array(joint_indices)
        self._n_params = int(sum(self._r_par_ind))
        self._r_tokens = tokens
        self._jc = JacobianCalculator(self._r_seq, self._r_par_ind, f_of_t=False,
                                      variables=self._r_var)
        if self.verbose:
            print("Went to base: " + ''.join(f"{t}({v})" for t, v in
                                             zip(self._to_base, self._to_base_vars)))
            print("Went to tool: " + ' '.join(f"{t}({v})" for t, v in
                                              zip(self._to_tool, self._to_tool_vars)))
        return self._r_seq, variables

    # --------------------------------------------------------- host -> kernel constants
    def _nominal_values(self, parameters):
        """Values of the parameter tokens: nominal + parameter (model_calibration.py:571-581)."""
        link = {str(p[0]): float(p[1]) for p in self._link_lengths} \
            if len(self._link_lengths) and not np.isscalar(self._link_lengths[0]) else {}
        return [link.get(str(nom), 0.0) + float(parameters[i])
                for i, nom in enumerate(self._r_nom_par)]

    def _compile(self, parameters=None, T_base=None, T_tools=None):
        """RkCalibModel for the current linearisation point."""
        if parameters is None:
            parameters = self._parameters
        pvals = self._nominal_values(parameters)
        Tb = np.eye(4) if T_base is None else np.array(T_base, dtype=np.float64)
        tools = [np.eye(4)] if not T_tools else [np.array(t, dtype=np.float64) for t in T_tools]
        # the same linearisation point again (evaluate after estimate, repeated evaluations): the
        # ~0.2 ms token walk below is skipped; the key holds every number the struct is built from
        key = (tuple(pvals), Tb.tobytes(), tuple(t.tobytes() for t in tools),
               np.asarray(self.qs_directions, dtype=np.float64).tobytes(),
               np.asarray(self.qs_offsets, dtype=np.float64).tobytes(), tuple(self._r_tokens), tuple(self._r_j_ind))
        cached = getattr(self, "_compiled", None)
        if cached is not None and cached[0] == key:
            return cached[1]
        m = _lib.RkCalibModel()
        toks, jind = self._r_tokens, self._r_j_ind
        if not jind or jind[0] != 1:
            raise ValueError("the reduced sequence must start with a joint token")
        nj = int(sum(jind))
        if nj > _lib.RK_CALIB_MAX_JOINTS or self._n_params > _lib.RK_CALIB_MAX_PARAMS:
            raise ValueError("model too large for the compiled kernel limits")
        m.njoints, m.nparams = nj, self._n_params
        j, p, X = -1, 0, np.eye(4)
        for tok, is_joint in zip(toks, jind):
            ax = "xyz".index(tok[1])
            if is_joint:
                if j >= 0:
                    _lib.fill(m.seg[j], _lib.xf12(X))
                j += 1
                X = np.eye(4)
                m.joint_axis[j] = ax
                m.joint_kind[j] = 1 if tok[0] == 'R' else 0
                m.joint_dir[j] = float(self.qs_directions[j])
                m.joint_off[j] = float(self.qs_offsets[j])
            else:
                m.param_joint[p] = j
                m.param_kind[p] = 1 if tok[0] == 'R' else 0
                _lib.fill(m.param_axis[p], X[:3, ax])
                _lib.fill(m.param_origin[p], X[:3, 3])
                X = X @ elementary_np(tok, pvals[p])
                p += 1
        _lib.fill(m.seg[j], _lib.xf12(X))
        _lib.fill(m.T_base, _lib.xf12(Tb))
        if len(tools) > _lib.RK_CALIB_MAX_TOOLS:
            raise ValueError("too many tools")
        m.ntools = len(tools)
        for t, Tt in enumerate(tools):
            _lib.fill(m.tool_p[t], Tt[:3, 3])
        self._compiled = (key, m)
        return m

    def _device_data(self, qs, pts, ntools=None):
        """(q [m, nj], pts [m, 3n]) as contiguous float64 CUDA tensors (uploads NumPy)."""
        def dev(x):
            if isinstance(x, torch.Tensor):
                _lib.require_cuda(x, "measurements")
                return x.to(torch.float64)
            return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64), device="cuda")
        q = dev(qs)
        p = dev(pts)
        p = p.reshape(p.shape[0], -1).contiguous()
        return q.contiguous(), p

    def _workspace(self, device):
        """Reduction workspace (CTA partials + ticket counter).  One per (device, stream): two
        streams of one device must not share the counter."""
        key = (str(device), torch.cuda.current_stream(device).cuda_stream)
        if key not in self._ws:
            nbytes = int(_lib.load().rk_calib_workspace_bytes())
            self._ws[key] = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        return self._ws[key]

    @staticmethod
    def _allreduce(t, op, process_group):
        import torch.distributed as dist
        if process_group is False or not (dist.is_available() and dist.is_initialized()):
            return
        group = None if process_group in (None, True) else process_group
        dist.all_reduce(t, op=op, group=group)

    # ------------------------------------------------------------------- reductions
    def normal_equations(self, qs, pts, T_base, T_tools, parameters=None, process_group=None):
        """H = sum J^T J [P,P], g = sum J^T r [P], rss = sum r^T r over this rank's
        measurements (all-reduced when torch.distributed is initialised).  CUDA tensors."""
        q, p = self._device_data(qs, pts)
        model = self._compile(parameters, T_base, T_tools)
        if p.shape[1] != 3 * model.ntools:
            raise ValueError("pts must hold 3 coordinates per tool")
        P = self._n_params
        out = torch.empty(P * (P + 1) // 2 + P + 1, dtype=torch.float64, device=q.device)
        ws = self._workspace(q.device)
        with torch.cuda.device(q.device):
            _lib.check(_lib.load().rk_calib_normal_eq_f64(
                C.byref(model), _lib.ptr(q), _lib.ptr(p), q.shape[0], _lib.ptr(out), _lib.ptr(ws),
                ws.numel(), _lib.stream_ptr(q.device)))
        import torch.distributed as dist
        self._allreduce(out, dist.ReduceOp.SUM if dist.is_available() else None, process_group)
        return out

    @staticmethod
    def unpack_normal_equations(out, P):
        """Packed upper triangle + g + rss (host NumPy) -> (H [P,P], g [P], rss)."""
        o = np.asarray(out, dtype=np.float64)
        H = np.zeros((P, P))
        iu = np.triu_indices(P)
        H[iu] = o[:P * (P + 1) // 2]
        H = H + np.triu(H, 1).T
        return H, o[P * (P + 1) // 2:P * (P + 1) // 2 + P].copy(), float(o[-1])

    def estimate_parameters(self, qs, pts, T_base=None, T_tools=None, process_group=None):
        """One Gauss-Newton style estimate of the geometric parameters
        (model_calibration.py:381-454).  qs (m, k), pts (m, 3n, 1) -- NumPy arrays are
        uploaded, CUDA tensors are used in place.  Returns the estimated parameters."""
        if self._parameters is None:
            print(self._no_model_msg)
            return [], []
        T_base = self._T_base if T_base is None else T_base
        T_tools = self._T_tools if T_tools is None else T_tools
        self._params_iteration += 1
        out = self.normal_equations(qs, pts, T_base, T_tools, process_group=process_group)
        J_sum, Jdp_sum, _ = self.unpack_normal_equations(out.cpu().numpy(), self._n_params)
        parameters = np.linalg.inv(J_sum).dot(Jdp_sum)              # :430
        simga = 0.043883037319796446
        Cov = simga**2 * J_sum
        s = self._step
        if self._params_iteration == 1:
            self._parameters = parameters
        else:
            self._parameters = (1.0 - s) * self._parameters + s * parameters
        if self.verbose:
            print()
            print("Parameters: ")
            for i, (var, val) in enumerate(zip(self._r_par, self._parameters)):
                print(f"{var} = {val:.4f} (+-{np.sqrt(Cov[i, i]):.4f})")
            print()
        self._last_normal_equations = (J_sum, Jdp_sum)
        return parameters

    def estimate_tool_base(self, qs, pts, parameters=None, process_group=None):
        """T_base and the n tool translations from the measurements
        (model_calibration.py:268-379).  Returns (T_base, [T_tool_i]) as sp.Matrix."""
        if self._parameters is None:
            print(self._no_model_msg)
            return [], []
        self._base_iteration += 1
        q, p = self._device_data(qs, pts)
        n = p.shape[1] // 3
        dim = 6 + 3 * n
        model = self._compile(self._parameters, None, [np.eye(4)] * n)
        out = torch.empty(dim * dim + dim, dtype=torch.float64, device=q.device)
        ws = self._workspace(q.device)
        with torch.cuda.device(q.device):
            _lib.check(_lib.load().rk_calib_tool_base_eq_f64(
                C.byref(model), _lib.ptr(q), _lib.ptr(p), q.shape[0], _lib.ptr(out), _lib.ptr(ws),
                ws.numel(), _lib.stream_ptr(q.device)))
        import torch.distributed as dist
        self._allreduce(out, dist.ReduceOp.SUM if dist.is_available() else None, process_group)
        o = out.cpu().numpy()
        A_sum, Adp_sum = o[:dim * dim].reshape(dim, dim), o[dim * dim:]
        estimated = np.linalg.inv(A_sum).dot(Adp_sum)               # :347
        s = self._step
        if self._base_iteration == 1:
            self._base_est = estimated
        else:
            self._base_est = (1 - s) * self._base_est + s * estimated
        self._last_tool_base_equations = (A_sum.copy(), Adp_sum.copy())
        return self._compose_base_tools(self._base_est)

    @staticmethod
    def _compose_base_tools(est):
        T_base = tf.get_Tx(est[0]) * tf.get_Ty(est[1]) * tf.get_Tz(est[2])
        T_base *= tf.get_Rx(est[3]) * tf.get_Ry(est[4]) * tf.get_Rz(est[5])
        T_tools = [tf.get_Tx(est[i]) * tf.get_Ty(est[i + 1]) * tf.get_Tz(est[i + 2])
                   for i in range(6, len(est), 3)]
        return T_base, T_tools

    def residual_sums(self, qs, pts, T_base, T_tools, parameters=None, process_group=None):
        """[sum dist^2, sum dx^2, sum dy^2, sum dz^2, max dist, max|dx|, max|dy|, max|dz|,
        number of measured points] (CUDA tensor) over this rank's measurements; with
        torch.distributed initialised the sums and the point count are all-reduced with SUM and the
        maxima with MAX, so every rank holds the statistics of the whole data set."""
        q, p = self._device_data(qs, pts)
        model = self._compile(parameters, T_base, T_tools)
        out = torch.empty(9, dtype=torch.float64, device=q.device)
        ws = self._workspace(q.device)
        with torch.cuda.device(q.device):
            _lib.check(_lib.load().rk_calib_eval_f64(
                C.byref(model), _lib.ptr(q), _lib.ptr(p), q.shape[0], _lib.ptr(out), _lib.ptr(ws),
                ws.numel(), _lib.stream_ptr(q.device)))
        out[8] = float(q.shape[0] * (p.shape[1] // 3))
        return self._reduce_residual_sums(out, process_group)

    def _reduce_residual_sums(self, out, process_group=None):
        """Cross-rank step of `residual_sums`: entries 0-3 and 8 (the sums and the point count)
        reduce with SUM, entries 4-7 (the maxima) with MAX."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and process_group is not False:
            sums, maxs = torch.cat([out[:4], out[8:]]), out[4:8].clone()
            self._allreduce(sums, dist.ReduceOp.SUM, process_group)
            self._allreduce(maxs, dist.ReduceOp.MAX, process_group)
            out = torch.cat([sums[:4], maxs, sums[4:]])
        return out

    @staticmethod
    def _metrics(o, n):
        """The eight numbers of evaluate (model_calibration.py:507-521) from the reduced sums."""
        return {
            "distance_rms": float(np.sqrt(o[0] / n)),
            "x_rms": float(np.sqrt(o[1] / n)),
            "y_rms": float(np.sqrt(o[2] / n)),
            "z_rms": float(np.sqrt(o[3] / n)),
            "max_dist": float(o[4]),
            "max_diff_x": float(o[5]),
            "max_diff_y": float(o[6]),
            "max_diff_z": float(o[7]),
        }

    def evaluate(self, qs, pts, T_base, T_tools, process_group=None, count=None):
        """RMS and maximum residuals (model_calibration.py:456-539).  When the measurements are
        sharded over the ranks of torch.distributed the residual sums AND the number of measured
        points are all-reduced, so the RMS values are those of the whole data set on every rank
        (`count` overrides the divisor, e.g. for replicated data with process_group=False)."""
        if self._parameters is None:
            print(self._no_model_msg)
            return [], []
        o = self.residual_sums(qs, pts, T_base, T_tools, process_group=process_group).cpu().numpy()
        result = self._metrics(o, count if count is not None else o[8])
        if self.verbose:
            print()
            print(f"Distance RMS (mm): {result['distance_rms']}")
            print(f"Max Distance diff (mm): {result['max_dist']}")
            for a in "xyz":
                print()
                print(f"{a.upper()} coordinate RMS (mm): {result[a + '_rms']}")
                print(f"Max {a.upper()} diff (mm): {result['max_diff_' + a]}")
            print()
        return result

    # ----------------------------------------------------------- single-pose helpers
    def _prepare_evaluation(self, qs, parameters=None):
        """Token values for one configuration (model_calibration.py:555-583)."""
        qs = np.multiply(qs, self.qs_directions) + self.qs_offsets
        if parameters is None:
            parameters = self._parameters
        pvals = self._nominal_values(parameters)
        res, ji, pi = [], 0, 0
        for is_joint in self._r_j_ind:
            if is_joint:
                res.append(qs[ji])
                ji += 1
            else:
                res.append(pvals[pi])
                pi += 1
        return res

    def get_pose(self, qs, T_base=None, T_tool=None, parameters=None):
        """T_base * chain * T_tool for one configuration (model_calibration.py:541-548)."""
        values = self._prepare_evaluation(qs, parameters)
        v = torch.tensor(np.asarray(values, dtype=np.float64).reshape(1, -1), device="cuda")
        T = chain_batch(self._r_tokens, v,
                        T_base=None if T_base is None else np.array(T_base, dtype=np.float64),
                        T_tool=None if T_tool is None else np.array(T_tool, dtype=np.float64),
                        layout="aos")
        return sp.Matrix(T[0].cpu().numpy())

    def get_jacobian(self, qs, T_base, T_tool):
        """3 x P position Jacobian w.r.t. the parameters (model_calibration.py:550-553)."""
        values = self._prepare_evaluation(qs)
        return self._jc.calculate_numerically(values, T_base, T_tool)[:3, :]

    # ------------------------------------------------------------------ optimisation
    def _cost(self, parameters, qs, pts, T_base, T_tools):
        return float(self.residual_sums(qs, pts, T_base, T_tools, parameters=parameters)[0])

    def optimize(self, qs, pts, T_base, T_tools):
        """scipy least_squares refinement (model_calibration.py:585-608); every cost
        evaluation is one GPU reduction instead of a Python loop over the measurements."""
        q, p = self._device_data(qs, pts)
        res = least_squares(lambda x: self._cost(x, q, p, T_base, T_tools), self._parameters,
                            jac='3-point', ftol=1e-4, gtol=1e-4, xtol=None,
                            verbose=2 if self.verbose else 0)
        self._parameters = res.x

    def optimize_with_base(self, qs, pts):
        """model_calibration.py:610-649."""
        q, p = self._device_data(qs, pts)
        n = len(self._parameters)

        def cost(x):
            T_base, T_tools = self._compose_base_tools(x[n:])
            return self._cost(x[:n], q, p, T_base, T_tools)

        res = least_squares(cost, np.concatenate((self._parameters, self._base_est)),
                            jac='3-point', ftol=1e-3, gtol=1e-3, xtol=1e-3,
                            verbose=2 if self.verbose else 0)
        if self.verbose:
            print(res)
        self._parameters = res.x
