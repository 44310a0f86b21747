"""SpeechRecognizer -- the reference's model object surface over the CUDA library.

Mirrors lvsr.bricks.recognizer.SpeechRecognizer (lvsr/bricks/recognizer.py:159-562):
same constructor keywords (``SpeechRecognizer(input_dims=..., input_num_chars=...,
eos_label=..., num_phonemes=..., name=..., data_prepend_eos=..., character_map=...,
**config['net'])``, lvsr/main.py:213-221), same method names and return conventions:

    initialize()                         parameters from the init schemes (reference: Blocks push/initialize)
    load_params(path) / save_params      Blocks checkpoint parameter naming
    cost(recordings, recordings_mask, labels, labels_mask)   -> costs [L, B]   (recognizer.py:375-390)
    analyze(inputs, groundtruth, prediction=None)            -> [costs[L], weights[L,T'], energies[L,T']]
    init_beam_search(beam_size); beam_search(inputs, **kw)   -> (outputs, costs)   (:496-533)

No symbolic graph exists: every method is a direct call into liblvsr_b200.so (C ABI in
include/lvsr_b200.h) on torch-owned device buffers.  There is no CPU fallback.
"""
import io
import logging
import tarfile
from collections import OrderedDict

import numpy as np

from . import _lib
from . import bricks as _bricks
from .search import BeamSearch, CandidateNotFoundError  # noqa: F401


logger = logging.getLogger(__name__)


class _Variable(object):
    """Stand-in for the Theano input variables lvsr/main.py reads the NAMES of (recognizer.inputs.keys(),
    recognizer.labels.name ...: lvsr/bricks/recognizer.py:351-361, lvsr/main.py:260-262,786)."""

    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return self.name


def _ptr(t):
    return None if t is None else t.data_ptr()


class _Child(object):
    """Named handle into the brick tree (``recognizer.generator.transition.attention`` ...)
    for code that walks the reference's attribute paths (lvsr/main.py:297-298,354-369)."""

    def __init__(self, name, **attrs):
        self.name = name
        self.children = []
        for k, v in attrs.items():
            setattr(self, k, v)


class SpeechRecognizer(object):
    def __init__(self, input_dims, input_num_chars, eos_label, num_phonemes,
                 dim_dec, dims_bidir, enc_transition=None, dec_transition=None,
                 use_states_for_readout=True, attention_type="content_and_conv",
                 criterion=None, bottom=None, lm=None, character_map=None,
                 bidir=True, subsample=None, dims_top=None, prior=None, conv_n=None,
                 post_merge_activation=None, post_merge_dims=None, dim_matcher=None,
                 embed_outputs=True, dim_output_embedding=None, dec_stack=1,
                 conv_num_filters=1, data_prepend_eos=True, energy_normalizer=None,
                 max_decoded_length_scale=1, name="recognizer", device=None, **kwargs):
        # ---- what the CUDA path implements; everything else fails loudly ----------
        def unsupported(what):
            raise NotImplementedError("attention-lvcsr_b200: %s is outside the B200 hot path "
                                      "(SURVEY.md section 8)" % what)
        if attention_type != "content_and_conv":
            unsupported("attention_type=%r" % attention_type)
        if lm:
            unsupported("language-model shallow fusion")
        if not bidir:
            unsupported("unidirectional encoder")
        if dims_top:
            unsupported("dims_top")
        if dec_stack != 1:
            unsupported("dec_stack > 1")
        if criterion is not None and criterion.get("name", "log_likelihood") != "log_likelihood":
            unsupported("criterion %r" % criterion.get("name"))
        if bottom and bottom.get("dims"):
            unsupported("bottom MLP")
        for tr in (enc_transition, dec_transition):
            if tr is not None and getattr(tr, "__name__", type(tr).__name__) != "GatedRecurrent":
                unsupported("transition %r" % tr)
        if post_merge_dims is not None and len(post_merge_dims) != 1:
            unsupported("deep post_merge")

        self.name = name
        self.eos_label = eos_label
        self.data_prepend_eos = data_prepend_eos
        self.character_map = character_map
        self.criterion = criterion or dict(name="log_likelihood")
        self.max_decoded_length_scale = max_decoded_length_scale
        self.rec_weights_init = None
        self.initial_states_init = None
        self.weights_init = None
        self.biases_init = None

        act = post_merge_activation if post_merge_activation is not None else _bricks.Tanh()
        if dim_matcher is None:
            dim_matcher = dim_dec                                  # recognizer.py:225-226
        if conv_n is None:
            raise ValueError("conv_n is required for content_and_conv attention")
        subsample = list(subsample) if subsample else [1] * len(dims_bidir)
        prior = dict(prior) if prior else dict(type="expanding", initial_begin=0, initial_end=10000,
                                               min_speed=0, max_speed=0)    # lvsr/bricks/attention.py:72-74
        self.net = dict(
            num_features=int(input_dims["recordings"]), dims_bidir=[int(d) for d in dims_bidir],
            subsample=[int(k) for k in subsample], dim_dec=int(dim_dec), dim_matcher=int(dim_matcher),
            conv_n=int(conv_n), conv_num_filters=int(conv_num_filters), num_phonemes=int(num_phonemes),
            # LookupFeedback(V+1, dim) or OneOfNFeedback(V+1) whose feedback is the one-hot vector (recognizer.py:278-284)
            dim_feedback=(int(dim_dec if dim_output_embedding is None else dim_output_embedding) if embed_outputs
                          else int(num_phonemes) + 1),
            embed_outputs=bool(embed_outputs),
            post_merge_dim=int(post_merge_dims[0]) if post_merge_dims else int(num_phonemes),
            post_merge_activation=act.kind, maxout_pieces=int(getattr(act, "num_pieces", 1)),
            use_states_for_readout=bool(use_states_for_readout),
            energy_normalizer=energy_normalizer or "softmax", prior=prior)
        if not post_merge_dims:
            # Readout's default post_merge is a bare Bias on readout_dim (sequence_generators.py:596-599)
            self.net["post_merge_activation"] = "identity"
            unsupported("readout without post_merge_dims")

        # brick-tree handles
        attention = _Child("conv_att", prior=prior, energy_normalizer=self.net["energy_normalizer"])
        transition = _Child("att_trans", attention=attention)
        readout = _Child("readout", emitter=_Child("emitter"), readout=None)
        self.generator = _Child("generator", transition=transition, readout=readout)
        self.encoder = _Child("encoder")
        self.top = _Child("top")
        self.bottom = _Child("bottom")
        self.children = [self.encoder, self.top, self.bottom, self.generator]

        # named inputs of the reference's graphs (lvsr/bricks/recognizer.py:351-361)
        self.inputs = OrderedDict(recordings=_Variable("recordings"))
        self.single_inputs = OrderedDict(recordings=_Variable("recordings"))
        self.inputs_mask = _Variable("recordings_mask")
        self.labels = _Variable("labels")
        self.labels_mask = _Variable("labels_mask")
        self.single_labels = _Variable("labels")
        self.n_steps = _Variable("n_steps")

        self._device = device
        self._handle = None
        self._beam_search = None
        self.beam_size = None
        self._ctor_state = None

    # ------------------------------------------------------------------------------
    # handle / device
    # ------------------------------------------------------------------------------
    def _torch(self):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("attention-lvcsr_b200 needs a CUDA device (no CPU fallback)")
        return torch

    @property
    def device(self):
        torch = self._torch()
        if self._device is None:
            self._device = torch.device("cuda", torch.cuda.current_device())
        return torch.device(self._device)

    def _make_config(self):
        import ctypes as C
        n = self.net
        cfg = _lib.LvsrConfig()
        cfg.num_features = n["num_features"]
        cfg.num_layers = len(n["dims_bidir"])
        for i, (d, k) in enumerate(zip(n["dims_bidir"], n["subsample"])):
            cfg.dims_bidir[i] = d
            cfg.subsample[i] = k
        cfg.dim_dec = n["dim_dec"]
        cfg.dim_matcher = n["dim_matcher"]
        cfg.conv_n = n["conv_n"]
        cfg.conv_num_filters = n["conv_num_filters"]
        cfg.num_phonemes = n["num_phonemes"]
        cfg.dim_feedback = n["dim_feedback"]
        cfg.post_merge_dim = n["post_merge_dim"]
        cfg.maxout_pieces = n["maxout_pieces"]
        cfg.post_merge_activation = _lib.ACTIVATIONS[n["post_merge_activation"]]
        cfg.use_states_for_readout = int(n["use_states_for_readout"])
        cfg.energy_normalizer = _lib.NORMALIZERS[n["energy_normalizer"]]
        p = n["prior"]
        cfg.prior_type = _lib.PRIORS[p.get("type", "expanding")]
        cfg.prior_initial_begin = float(p.get("initial_begin", 0))
        cfg.prior_initial_end = float(p.get("initial_end", 10000))
        cfg.prior_min_speed = float(p.get("min_speed", 0))
        cfg.prior_max_speed = float(p.get("max_speed", 0))
        cfg.prior_before = float(p.get("before", 0))
        cfg.prior_after = float(p.get("after", 0))
        cfg.one_of_n_feedback = 0 if n.get("embed_outputs", True) else 1
        return cfg

    def _require_ready(self):
        if self._handle is None:
            import ctypes as C
            torch = self._torch()
            lib = _lib.load()
            with torch.cuda.device(self.device):
                h = C.c_void_p()
                cfg = self._make_config()
                _lib.check(lib.lvsr_model_create(C.byref(cfg), C.byref(h)))
            self._handle = h
        return self._handle

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.load().lvsr_model_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    def _stream(self):
        return self._torch().cuda.current_stream(self.device).cuda_stream

    # ------------------------------------------------------------------------------
    # parameters
    # ------------------------------------------------------------------------------
    def parameter_shapes(self):
        import ctypes as C
        lib, h = _lib.load(), self._require_ready()
        out = OrderedDict()
        for i in range(lib.lvsr_model_num_params(h)):
            shape = (C.c_int64 * 2)()
            ndim = C.c_int32()
            _lib.check(lib.lvsr_model_param_shape(h, i, shape, C.byref(ndim)))
            out[lib.lvsr_model_param_name(h, i).decode()] = tuple(int(shape[j]) for j in range(ndim.value))
        return out

    def set_parameter_values(self, values):
        """Model.set_parameter_values: {Blocks parameter path: ndarray}."""
        lib, h = _lib.load(), self._require_ready()
        shapes = self.parameter_shapes()
        for name, value in values.items():
            if name not in shapes:
                raise KeyError("unknown parameter %s" % name)
            arr = np.ascontiguousarray(value, dtype=np.float32)
            if tuple(arr.shape) != shapes[name]:
                raise ValueError("parameter %s: expected shape %s, got %s" % (name, shapes[name], arr.shape))
            _lib.check(lib.lvsr_model_set_param(h, name.encode(), arr.ctypes.data, arr.size))
        _lib.check(lib.lvsr_model_finalize(h))

    def get_parameter_values(self):
        lib, h = _lib.load(), self._require_ready()
        out = OrderedDict()
        for name, shape in self.parameter_shapes().items():
            arr = np.empty(shape, dtype=np.float32)
            _lib.check(lib.lvsr_model_get_param(h, name.encode(), arr.ctypes.data, arr.size))
            out[name] = arr
        return out

    def initialize(self, seed=1):
        """Blocks ``initialize()``: one RandomState walked in brick order; a recurrent
        brick takes rec_weights_init for all three matrices, initial states take
        initial_states_init (lvsr/bricks/recognizer.py:363-373; B/bricks/recurrent.py:568-580)."""
        w_init = self.weights_init or _bricks.IsotropicGaussian(0.01)
        b_init = self.biases_init or _bricks.Constant(0.0)
        rec_init = self.rec_weights_init or w_init
        h0_init = self.initial_states_init or _bricks.Constant(0.0)
        rng = np.random.RandomState(seed)
        values = OrderedDict()
        for name, shape in self.parameter_shapes().items():
            leaf = name.rsplit(".", 1)[1]
            if leaf == "b":
                v = b_init.generate(rng, shape)
            elif leaf == "state_to_state":
                v = rec_init.generate(rng, shape)
            elif leaf == "state_to_gates":
                d = shape[0]
                v = np.hstack([rec_init.generate(rng, (d, d)), rec_init.generate(rng, (d, d))])
            elif leaf == "initial_state":
                v = h0_init.generate(rng, shape)
            else:
                v = w_init.generate(rng, shape)
            values[name] = np.asarray(v, dtype=np.float32).reshape(shape)
        self.set_parameter_values(values)

    def load_params(self, path):
        """Blocks checkpoint (tar with a ``_parameters`` npz whose keys use '|' for '/':
        libs/blocks/blocks/serialization.py:264-282,606-610) or a plain .npz.  Like
        Model.set_parameter_values (libs/blocks/blocks/model.py:120-146) unknown names and missing parameters
        are LOGGED, not raised; missing parameters keep their current values."""
        data = None
        if tarfile.is_tarfile(path):
            with tarfile.open(path) as tar:
                data = np.load(io.BytesIO(tar.extractfile("_parameters").read()))
        else:
            data = np.load(path)
        values = {k.replace("|", "/"): data[k] for k in data.files}
        shapes = self.parameter_shapes()
        unknown = sorted(set(values) - set(shapes))
        missing = sorted(set(shapes) - set(values))
        if unknown:
            logger.error("unknown parameter names: {}\n".format(unknown))
        if missing:
            logger.error("missing values for parameters: {}\n".format(missing))
        self.set_parameter_values({k: v for k, v in values.items() if k in shapes})
        return dict(unknown=unknown, missing=missing)

    def save_params(self, path):
        """Write the parameters the way blocks.serialization.dump stores them separately: a tar archive with one
        member ``_parameters`` = numpy.savez of {brick path with '|' for '/': array}
        (libs/blocks/blocks/serialization.py:136,264-282,493-500,606-610) -- readable by the reference's
        load_parameters and by load_params above."""
        buf = io.BytesIO()
        np.savez(buf, **{k.replace("/", "|"): v for k, v in self.get_parameter_values().items()})
        payload = buf.getvalue()
        with tarfile.open(path, "w") as tar:
            info = tarfile.TarInfo("_parameters")
            info.size = len(payload)
            tar.addfile(info, io.BytesIO(payload))

    # pickling: device handles do not travel (lvsr/bricks/recognizer.py:549-562 drops the compiled functions)
    def __getstate__(self):
        state = dict(self.__dict__)
        for attr in ("_handle", "_beam_search", "_generator_state"):
            state.pop(attr, None)
        state["_device"] = None if self._device is None else str(self._device)
        state["_saved_parameters"] = None if self._handle is None else self.get_parameter_values()
        return state

    def __setstate__(self, state):
        saved = state.pop("_saved_parameters", None)
        self.__dict__.update(state)
        self._handle = None
        self._beam_search = None
        if saved is not None:
            try:
                self.set_parameter_values(saved)
            except RuntimeError:           # unpickled where no GPU is visible: parameters stay on the host copy
                self._pending_parameters = saved

    # ------------------------------------------------------------------------------
    # device-side operators (torch tensors in, torch tensors out)
    # ------------------------------------------------------------------------------
    def _dev(self, a, dtype=None):
        torch = self._torch()
        if a is None:
            return None
        if isinstance(a, torch.Tensor):
            t = a.to(self.device)
        else:
            t = torch.as_tensor(np.ascontiguousarray(a), device=self.device)
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        return t.contiguous()

    def launch_status(self):
        """(status, stepwise_fallbacks) of the persistent decoder (lvsr_model_status): status 0 = the last
        cost_matrix launch completed; the counter says how often a failed launch was re-run step-wise."""
        import ctypes as C
        lib, h = _lib.load(), self._require_ready()
        st, fb = C.c_int32(), C.c_int64()
        _lib.check(lib.lvsr_model_status(h, C.byref(st), C.byref(fb)))
        return int(st.value), int(fb.value)

    def encoded_length(self, T):
        return int(_lib.load().lvsr_encoded_length(self._require_ready(), int(T)))

    @property
    def dim_encoded(self):
        return 2 * self.net["dims_bidir"][-1]

    def encode(self, recordings, recordings_mask=None):
        """Encoder.apply (lvsr/bricks/__init__.py:71-78): [T,B,F], [T,B] -> ([T',B,E], [T',B])."""
        torch = self._torch()
        lib, h = _lib.load(), self._require_ready()
        x = self._dev(recordings, torch.float32)
        m = self._dev(recordings_mask, torch.float32)
        if x.dim() != 3:
            raise ValueError("encode: recordings [T,B,F] expected")
        T, B, F = x.shape
        if F != self.net["num_features"]:
            raise ValueError("expected %d features, got %d" % (self.net["num_features"], F))
        if m is not None and tuple(m.shape) != (T, B):
            raise ValueError("encode: recordings_mask must be [%d, %d], got %s" % (T, B, tuple(m.shape)))
        Tp = self.encoded_length(T)
        att = torch.empty((Tp, B, self.dim_encoded), dtype=torch.float32, device=self.device)
        attm = torch.empty((Tp, B), dtype=torch.float32, device=self.device)
        _lib.check(lib.lvsr_encoder_forward(h, _ptr(x), _ptr(m), T, B, _ptr(att), _ptr(attm), self._stream()))
        return att, attm

    def preprocess(self, attended):
        torch = self._torch()
        lib, h = _lib.load(), self._require_ready()
        Tp, U, _ = attended.shape
        out = torch.empty((Tp, U, self.net["dim_matcher"]), dtype=torch.float32, device=self.device)
        _lib.check(lib.lvsr_preprocess(h, _ptr(attended), Tp, U, _ptr(out), self._stream()))
        return out

    def _check_labels(self, labels):
        """Theano's lookup raises IndexError on a symbol outside the table; host arrays are
        checked here (device tensors are the caller's contract: no hidden synchronisation)."""
        if isinstance(labels, np.ndarray) and labels.size:
            lo, hi = int(labels.min()), int(labels.max())
            if lo < 0 or hi >= self.net["num_phonemes"]:
                raise ValueError("labels must lie in [0, %d): got %d..%d" % (self.net["num_phonemes"], lo, hi))

    def cost_matrix(self, labels, labels_mask, attended, attended_mask, return_all=False):
        """generator.cost_matrix (B/bricks/sequence_generators.py:319-326) on device tensors."""
        torch = self._torch()
        lib, h = _lib.load(), self._require_ready()
        self._check_labels(labels)
        y = self._dev(labels, torch.int64)
        ym = self._dev(labels_mask, torch.float32)
        att = self._dev(attended, torch.float32)
        attm = self._dev(attended_mask, torch.float32)
        if y.dim() != 2 or att.dim() != 3 or attm.dim() != 2:
            raise ValueError("cost_matrix: labels [L,B], attended [T',B,E], attended_mask [T',B] expected")
        L, B = y.shape
        Tp = att.shape[0]
        if L < 1 or B < 1 or Tp < 1:
            raise ValueError("cost_matrix: empty labels or attended sequence")
        if tuple(att.shape) != (Tp, B, self.dim_encoded):
            raise ValueError("cost_matrix: attended must be [%d, %d, %d], got %s" % (Tp, B, self.dim_encoded, tuple(att.shape)))
        if tuple(attm.shape) != (Tp, B):
            raise ValueError("cost_matrix: attended_mask must be [%d, %d], got %s" % (Tp, B, tuple(attm.shape)))
        if ym is not None and tuple(ym.shape) != (L, B):
            raise ValueError("cost_matrix: labels_mask must be [%d, %d], got %s" % (L, B, tuple(ym.shape)))
        costs = torch.empty((L, B), dtype=torch.float32, device=self.device)
        extra = {}
        if return_all:
            extra = dict(weights=torch.empty((L, B, Tp), dtype=torch.float32, device=self.device),
                         energies=torch.empty((L, B, Tp), dtype=torch.float32, device=self.device),
                         states=torch.empty((L, B, self.net["dim_dec"]), dtype=torch.float32, device=self.device),
                         weighted_averages=torch.empty((L, B, self.dim_encoded), dtype=torch.float32,
                                                       device=self.device))
        _lib.check(lib.lvsr_cost_matrix(
            h, _ptr(att), _ptr(attm), Tp, B, _ptr(y), _ptr(ym), L, _ptr(costs),
            _ptr(extra.get("weights")), _ptr(extra.get("energies")), _ptr(extra.get("states")),
            _ptr(extra.get("weighted_averages")), self._stream()))
        if return_all:
            extra["costs"] = costs
            return extra
        return costs

    # ------------------------------------------------------------------------------
    # reference-facing methods (numpy in, numpy out)
    # ------------------------------------------------------------------------------
    def cost(self, recordings, recordings_mask, labels, labels_mask):
        """SpeechRecognizer.cost (recognizer.py:375-390) through the host-buffer C entry
        point: copies in, encoder + teacher-forced decoder, costs [L, B] copied out."""
        lib, h = _lib.load(), self._require_ready()
        x = np.ascontiguousarray(recordings, dtype=np.float32)
        m = None if recordings_mask is None else np.ascontiguousarray(recordings_mask, dtype=np.float32)
        y = np.ascontiguousarray(labels, dtype=np.int64)
        self._check_labels(y)
        ym = None if labels_mask is None else np.ascontiguousarray(labels_mask, dtype=np.float32)
        if x.ndim != 3 or y.ndim != 2:
            raise ValueError("cost: recordings [T,B,F] and labels [L,B] expected")
        T, B, F = x.shape
        L = y.shape[0]
        if F != self.net["num_features"]:
            raise ValueError("expected %d features, got %d" % (self.net["num_features"], F))
        if T < 1 or B < 1 or L < 1:
            raise ValueError("cost: empty batch")
        if m is not None and m.shape != (T, B):
            raise ValueError("cost: recordings_mask must be [%d, %d], got %s" % (T, B, m.shape))
        if y.shape != (L, B):
            raise ValueError("cost: labels must be [L, %d], got %s" % (B, y.shape))
        if ym is not None and ym.shape != (L, B):
            raise ValueError("cost: labels_mask must be [%d, %d], got %s" % (L, B, ym.shape))
        costs = np.empty((L, B), dtype=np.float32)
        torch = self._torch()
        with torch.cuda.device(self.device):
            _lib.check(lib.lvsr_recognizer_cost_host(
                h, x.ctypes.data, None if m is None else m.ctypes.data, y.ctypes.data,
                None if ym is None else ym.ctypes.data, T, B, L, costs.ctypes.data, self._stream()))
        return costs

    def analyze(self, inputs, groundtruth, prediction=None):
        """recognizer.py:452-494: one utterance, mask of ones, no label mask."""
        rec = np.asarray(dict(inputs)["recordings"], dtype=np.float32)[:, None, :]
        labels = np.asarray(groundtruth if prediction is None else prediction, dtype=np.int64)[:, None]
        att, attm = self.encode(rec, np.ones(rec.shape[:2], dtype=np.float32))
        r = self.cost_matrix(labels, None, att, attm, return_all=True)
        return [r["costs"][:, 0].cpu().numpy(), r["weights"][:, 0, :].cpu().numpy(),
                r["energies"][:, 0, :].cpu().numpy()]

    def init_beam_search(self, beam_size):
        """recognizer.py:496-511."""
        if self._beam_search is not None and self.beam_size == beam_size:
            return
        self.beam_size = beam_size
        self._beam_search = BeamSearch(beam_size, self)
        self._beam_search.compile()

    def beam_search(self, inputs, **kwargs):
        """recognizer.py:513-533: inputs {'recordings': [T, F]} -> (outputs, costs)."""
        self.init_beam_search(self.beam_size)
        inputs = dict(inputs)
        rec = np.asarray(inputs.pop("recordings"), dtype=np.float32)
        if inputs:
            raise Exception("Unknown inputs passed to beam search: {}".format(list(inputs.keys())))
        max_length = int(rec.shape[0] / self.max_decoded_length_scale)
        return self._beam_search.search({"recordings": rec[:, None, :]}, self.eos_label, max_length,
                                        ignore_first_eol=self.data_prepend_eos, **kwargs)

    def beam_search_many(self, inputs_list, **kwargs):
        """beam_search for a list of {'recordings': [T_u, F]} decoded together on the GPU (one set of launches per
        step for all utterances, BeamSearch.search_many); returns [(outputs, costs), ...] in order.  The
        reference decodes one utterance at a time (lvsr/main.py:806-821 loops over the data stream)."""
        self.init_beam_search(self.beam_size)
        recs = []
        for inputs in inputs_list:
            inputs = dict(inputs)
            recs.append(np.asarray(inputs.pop("recordings"), dtype=np.float32))
            if inputs:
                raise Exception("Unknown inputs passed to beam search: {}".format(list(inputs.keys())))
        max_lengths = [int(x.shape[0] / self.max_decoded_length_scale) for x in recs]
        return self._beam_search.search_many(recs, self.eos_label, max_lengths,
                                             ignore_first_eol=self.data_prepend_eos, **kwargs)

    # ---- generate / sample (B/bricks/sequence_generators.py:328-377; recognizer.py:535-547) ----
    def generate(self, recordings, recordings_mask=None, n_steps=None, sample=True, seed=None):
        """SequenceGenerator.generate iterated n_steps times for a batch [T,B,F]: glimpses -> readout -> emit ->
        feedback -> next state.  ``sample=True`` emits from the softmax like SoftmaxEmitter.emit
        (sequence_generators.py:772-778; a seeded Philox stream on the device instead of Theano's MRG stream, so
        draws differ from the reference while their distribution does not), ``sample=False`` emits the arg-max.
        Returns dict(outputs [n,B] int64, costs [n,B] = -log p(emitted), states [n,B,C], weights [n,B,T'])."""
        torch = self._torch()
        att, attm = self.encode(recordings, recordings_mask)
        B, Tp = att.shape[1], att.shape[0]
        if n_steps is None:
            n_steps = int(np.asarray(recordings).shape[0] / self.max_decoded_length_scale)
        ctx = dict(attended=att, attended_mask=attm, preprocessed=self.preprocess(att))
        st = self._initial_states(Tp, B)
        gen = None
        if sample:
            gen = torch.Generator(device=self.device)
            gen.manual_seed(1 if seed is None else int(seed))
        outs, costs, states, weights = [], [], [], []
        for _ in range(int(n_steps)):
            neglogp = self._logprobs(ctx, st)
            if sample:
                y = torch.multinomial(torch.exp(-neglogp), 1, generator=gen)[:, 0]
            else:
                y = neglogp.argmin(dim=1)
            costs.append(neglogp.gather(1, y[:, None])[:, 0])
            st = self._next_states(ctx, st, y)
            outs.append(y)
            states.append(st["states"])
            weights.append(st["weights"])
        return dict(outputs=torch.stack(outs).cpu().numpy(), costs=torch.stack(costs).cpu().numpy(),
                    states=torch.stack(states).cpu().numpy(), weights=torch.stack(weights).cpu().numpy())

    def sample(self, inputs, n_steps=None, seed=None):
        """recognizer.py:540-547: one utterance {'recordings': [T,F]} -> sampled outputs [n_steps, 1]."""
        rec = np.asarray(dict(inputs)["recordings"], dtype=np.float32)[:, None, :]
        if n_steps is None:
            n_steps = int(rec.shape[0] / self.max_decoded_length_scale)
        return self.generate(rec, None, n_steps=n_steps, sample=True, seed=seed)["outputs"]

    def get_generate_graph(self, use_mask=True, n_steps=None):
        """recognizer.py:414-421 returns the symbolic generate application; here: a callable with the same inputs
        (recordings [, recordings_mask], n_steps) returning the dict of generate()."""
        def run(recordings, recordings_mask=None, n_steps=n_steps, **kw):
            return self.generate(recordings, recordings_mask if use_mask else None, n_steps=n_steps, **kw)
        return run

    def get_cost_graph(self, batch=True, prediction=None, prediction_mask=None):
        """recognizer.py:423-450: the cost 'graph' as a callable: batch=True takes (recordings, recordings_mask, labels,
        labels_mask) -> costs [L,B]; batch=False takes one utterance (recordings [T,F], labels [L]) -> costs [L]."""
        if batch:
            return lambda recordings, recordings_mask, labels, labels_mask: self.cost(recordings, recordings_mask, labels, labels_mask)
        return lambda recordings, labels: self.analyze({"recordings": recordings}, labels)[0]

    # ------------------------------------------------------------------------------
    # BeamSearch state functions (C-ABI calls)
    # ------------------------------------------------------------------------------
    def _initial_states(self, Tp, R):
        torch = self._torch()
        lib, h = _lib.load(), self._require_ready()
        dev = self.device
        st = OrderedDict(
            states=torch.empty((R, self.net["dim_dec"]), dtype=torch.float32, device=dev),
            outputs=torch.empty((R,), dtype=torch.int64, device=dev),
            weighted_averages=torch.empty((R, self.dim_encoded), dtype=torch.float32, device=dev),
            weights=torch.empty((R, Tp), dtype=torch.float32, device=dev),
            energies=torch.empty((R, Tp), dtype=torch.float32, device=dev),
            step=torch.empty((R,), dtype=torch.int64, device=dev))
        _lib.check(lib.lvsr_initial_states(h, Tp, R, _ptr(st["states"]), _ptr(st["outputs"]),
                                           _ptr(st["weighted_averages"]), _ptr(st["weights"]),
                                           _ptr(st["energies"]), _ptr(st["step"]), self._stream()))
        return st

    def _row_utt(self, contexts, R):
        torch = self._torch()
        U = contexts["attended"].shape[1]
        ru = contexts.get("row_utt")
        if ru is None:
            if U == R:
                return None
            if U != 1:
                raise ValueError("contexts hold %d utterances for %d rows: pass row_utt" % (U, R))
            ru = torch.zeros((R,), dtype=torch.int32, device=self.device)
        return self._dev(ru, torch.int32)

    def _logprobs(self, contexts, st):
        torch = self._torch()
        lib, h = _lib.load(), self._require_ready()
        att = contexts["attended"]
        Tp, U, _ = att.shape
        R = st["states"].shape[0]
        ru = self._row_utt(contexts, R)
        out = torch.empty((R, self.net["num_phonemes"]), dtype=torch.float32, device=self.device)
        _lib.check(lib.lvsr_logprobs(h, _ptr(att), _ptr(contexts.get("preprocessed")), _ptr(contexts["attended_mask"]),
                                     Tp, U, _ptr(ru), R, _ptr(st["states"].contiguous()),
                                     _ptr(st["weights"].contiguous()), _ptr(st["step"].contiguous()), _ptr(out),
                                     self._stream()))
        return out

    def _next_states(self, contexts, st, outputs):
        torch = self._torch()
        lib, h = _lib.load(), self._require_ready()
        att = contexts["attended"]
        Tp, U, _ = att.shape
        R = st["states"].shape[0]
        ru = self._row_utt(contexts, R)
        y = self._dev(outputs, torch.int64)
        nxt = OrderedDict(
            states=torch.empty_like(st["states"]), outputs=y,
            weighted_averages=torch.empty((R, self.dim_encoded), dtype=torch.float32, device=self.device),
            weights=torch.empty((R, Tp), dtype=torch.float32, device=self.device),
            energies=torch.empty((R, Tp), dtype=torch.float32, device=self.device),
            step=torch.empty((R,), dtype=torch.int64, device=self.device))
        _lib.check(lib.lvsr_next_states(
            h, _ptr(att), _ptr(contexts.get("preprocessed")), _ptr(contexts["attended_mask"]), Tp, U, _ptr(ru), R,
            _ptr(st["states"].contiguous()), _ptr(st["weights"].contiguous()), _ptr(st["step"].contiguous()),
            _ptr(y), _ptr(nxt["states"]), _ptr(nxt["weighted_averages"]), _ptr(nxt["weights"]),
            _ptr(nxt["energies"]), _ptr(nxt["step"]), self._stream()))
        return nxt
