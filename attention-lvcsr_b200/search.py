"""Beam search over the CUDA decoder -- mirror of blocks.search.BeamSearch.

Same constructor/driver contract as the reference's (modified) class
(libs/blocks/blocks/search.py:19-407): ``BeamSearch(beam_size, recognizer)``,
``compile()``, ``search(input_values, eol_symbol, max_length, ...)`` returning
``(outputs, costs)``.  The four compiled Theano functions become four C-ABI calls
(lvsr_encoder_forward, lvsr_initial_states, lvsr_logprobs, lvsr_next_states); all
hypothesis state stays on the GPU between steps, only the [width, V] cost table
crosses to the host for the k-best selection.  Differences from the reference that
do not change results: the encoded sequence is NOT replicated per hypothesis (rows
index their utterance) and attention.preprocess runs once per utterance instead of
twice per step.
"""
import numpy as np

from . import _lib


class CandidateNotFoundError(Exception):
    """libs/blocks/blocks/search.py:15-16."""


def _smallest(matrix, k):
    """k smallest entries of a matrix: ((rows, cols), values), increasing
    (libs/blocks/blocks/search.py:220-242; numpy's argpartition/argsort tie order)."""
    flat = matrix.reshape(-1)
    if flat.shape[0] > k:
        keep = np.argpartition(flat, k)[:k]
    else:
        keep = np.arange(flat.shape[0])
    keep = keep[np.argsort(flat[keep])]
    return np.unravel_index(keep, matrix.shape), flat[keep]


class BeamSearch(object):
    def __init__(self, beam_size, recognizer):
        self.beam_size = beam_size
        self.recognizer = recognizer
        self.compiled = False
        self.context_names = ["attended", "attended_mask"]
        self.state_names = ["states", "outputs", "weighted_averages", "weights", "energies", "step"]

    _smallest = staticmethod(_smallest)

    def compile(self):
        """Nothing to compile: the kernels are ahead-of-time sm_100a code."""
        self.recognizer._require_ready()
        self.compiled = True

    # ---- the four device functions ---------------------------------------------
    def compute_contexts(self, recordings):
        """recordings [T, 1, F] (numpy or torch) -> dict(attended, attended_mask, preprocessed)."""
        r = self.recognizer
        att, mask = r.encode(recordings, None)
        return dict(attended=att, attended_mask=mask, preprocessed=r.preprocess(att))

    def compute_initial_states(self, contexts, width=1):
        return self.recognizer._initial_states(contexts["attended"].shape[0], width)

    def compute_logprobs(self, contexts, states):
        return self.recognizer._logprobs(contexts, states)

    def compute_next_states(self, contexts, states, outputs):
        return self.recognizer._next_states(contexts, states, outputs)

    # ---- driver -----------------------------------------------------------------
    def search(self, input_values, eol_symbol, max_length, ignore_first_eol=False, as_arrays=False,
               char_discount=0, round_to_inf=1e9, stop_on="patience", validate_solution_function=None):
        """See the reference docstring (libs/blocks/blocks/search.py:244-288).
        ``input_values``: {'recordings': array [T, 1, F]} (name or any single key)."""
        import torch
        if not self.compiled:
            self.compile()
        (recordings,) = list(input_values.values())
        contexts = self.compute_contexts(recordings)
        states = self.compute_initial_states(contexts)

        outputs_hist = states["outputs"].cpu().numpy()[None, :]       # includes the initial symbol
        costs_hist = np.zeros(outputs_hist.shape, dtype=np.float32)
        done = []
        min_cost = 1000
        patience = None

        def discounted(item):
            return item[1][-1] - char_discount * len(item[1])

        for i in range(max_length):
            if states["states"].shape[0] == 0:
                break
            if stop_on == "patience":
                done = sorted(done, key=discounted)[:self.beam_size]
                if done:
                    best = discounted(done[0])
                    if best < min_cost:
                        min_cost, patience = best, 30
                    else:
                        patience -= 1
                        if patience == 0:
                            break
            elif stop_on == "optimistic_future_cost":
                if len(done) >= self.beam_size:
                    optimistic = costs_hist[-1, :].min() - char_discount * max_length
                    last = done[self.beam_size - 1][1]
                    if last[-1] - char_discount * len(last) < optimistic:
                        break
            else:
                raise ValueError("Unknown stopping criterion {}".format(stop_on))

            logprobs = self.compute_logprobs(contexts, states).cpu().numpy()
            assert np.isfinite(logprobs).all()
            next_costs = costs_hist[-1, :, None] + logprobs
            (parents, symbols), chosen = self._smallest(next_costs, self.beam_size)

            sel = torch.as_tensor(parents, device=states["states"].device)
            states = {k: v.index_select(0, sel) for k, v in states.items()}
            outputs_hist = np.take(outputs_hist, parents, axis=1)
            costs_hist = np.take(costs_hist, parents, axis=1)

            states = self.compute_next_states(contexts, states, symbols)
            outputs_hist = np.vstack([outputs_hist, symbols[None, :]])
            costs_hist = np.vstack([costs_hist, chosen[None, :].astype(costs_hist.dtype)])

            alive = symbols != eol_symbol
            if ignore_first_eol and i == 0:
                alive[:] = True
            ended = np.where((outputs_hist[-1] == eol_symbol) &
                             (costs_hist[-1] - costs_hist[-2] < round_to_inf))[0]
            for idx in ended:
                if (validate_solution_function is None or
                        validate_solution_function(input_values, outputs_hist[:, idx])):
                    done.append((outputs_hist[:, idx], costs_hist[:, idx]))
            keep = np.where(alive)[0]
            sel = torch.as_tensor(keep, device=states["states"].device)
            states = {k: v.index_select(0, sel) for k, v in states.items()}
            outputs_hist = np.take(outputs_hist, keep, axis=1)
            costs_hist = np.take(costs_hist, keep, axis=1)

        if not done:
            raise CandidateNotFoundError()
        done = sorted(done, key=discounted)

        max_len = max(seq.shape[0] for seq, _ in done)
        all_outputs = np.zeros((max_len, len(done)))
        all_masks = np.zeros((max_len, len(done)))
        all_costs = np.zeros((max_len, len(done)))
        for j, (seq, cost) in enumerate(done):
            all_outputs[:len(seq), j] = seq
            all_masks[:len(seq), j] = 1
            all_costs[:len(cost), j] = cost
            all_costs[len(cost):, j] = cost[-1]
        result = (all_outputs[1:], all_masks[1:], all_costs[1:] - all_costs[:-1])
        if as_arrays:
            return result
        return self.result_to_lists(result)

    @staticmethod
    def result_to_lists(result):
        outputs, masks, costs = [a.T for a in result]
        outputs = [[int(t) for t in out[:int(m.sum())]] for out, m in zip(outputs, masks)]
        costs = [float(c) for c in costs.T.sum(axis=0)]
        return outputs, costs
