"""Drop-in counterparts of ``bpreveal.models`` (src/models.py of the reference).

The three builders keep the reference's names, argument order and meaning
(``soloModel`` models.py:53-56, ``transformationModel`` :226-229, ``combinedModel`` :268-272) but
return model objects backed by the B200 engine (``engine.SoloNet``) instead of Keras graphs.  The
objects expose the part of the Keras ``Model`` duck-type that the reference's tools touch
(SURVEY.md 8b-2): ``name``, ``inputs`` / ``input`` / ``outputs`` / ``output`` with ``.shape``,
``trainable``, ``compile``, ``fit``, ``predict``, ``save``, ``summary``, ``optimizer.learning_rate``,
``get_weights`` / ``set_weights`` in Keras layer order, and the reference's layer names
(``{name}_initial_conv``, ``{name}_conv_{i}``, ``solo_profile_{head}``, ``solo_logcounts_{head}``,
``regress_linear_*_conv``, ... models.py:39-48, 85-104, 141-172; layers.py:35-46).
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from . import ops
from .engine import NetConfig, SoloNet

NUM_BASES = 4  # internal/constants.py:152


class _TensorSpec:
    """Stand-in for a symbolic Keras tensor: only ``.shape`` and ``.name`` are consulted
    (utils.py:881-887, models.py:339)."""

    def __init__(self, shape, name):
        self.shape = tuple(shape)
        self.name = name

    def __repr__(self):
        return f"<TensorSpec {self.name} shape={self.shape}>"


class _Optimizer:
    """``keras.optimizers.Adam`` stand-in: the tools read and assign ``learning_rate``."""

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.learning_rate = float(learning_rate)
        self.beta_1, self.beta_2, self.epsilon = beta_1, beta_2, epsilon


def Adam(learning_rate=0.001, **kw):  # noqa: N802  (Keras spelling)
    return _Optimizer(learning_rate, **kw)


class _LazyBatchLogs:
    """The ``logs`` argument of the per-batch callbacks: metric names are known up front, values
    (running means of the epoch so far) are read back from the device on first access."""

    def __init__(self, model, sums, sums_dev, nseq):
        self._model, self._sums, self._dev, self._n = model, sums, sums_dev, nseq
        self._vals = None

    def _materialise(self):
        if self._vals is None:
            tot = self._sums.copy()
            if self._dev is not None:
                tot = tot + self._dev.cpu().numpy()
            self._vals = self._model._epoch_logs(tot / max(self._n, 1), "")
        return self._vals

    def keys(self):
        if self._vals is None:  # names only: no synchronisation
            H = len(self._model.headTasks)
            return list(self._model._epoch_logs(np.zeros(2 * H), "").keys())
        return list(self._vals.keys())

    def __iter__(self):
        return iter(self.keys())

    def __contains__(self, k):
        return k in self.keys()

    def __getitem__(self, k):
        return self._materialise()[k]

    def get(self, k, default=None):
        return self._materialise().get(k, default)

    def items(self):
        return self._materialise().items()


class History:
    """``keras.callbacks.History``: ``.history`` maps metric name -> list over epochs."""

    def __init__(self):
        self.history = {}
        self.epoch = []

    def append(self, epoch, logs):
        self.epoch.append(epoch)
        for k, v in logs.items():
            self.history.setdefault(k, []).append(float(v))


def _device(device=None):
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        raise RuntimeError("bpreveal_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _head_fields(headList):
    names = [h["head-name"] for h in headList]
    tasks = [int(h["num-tasks"]) for h in headList]
    return names, tasks


class Model:
    """Behaviour shared by the three model kinds."""

    useOldKeras = False  # utils.py:83 looks for this attribute

    def __init__(self, name, inputLength, outputLength, headList):
        self.name = name
        self.headList = headList
        self.headNames, self.headTasks = _head_fields(headList)
        self.inputLength, self.outputLength = inputLength, outputLength
        self.inputs = [_TensorSpec((None, inputLength, NUM_BASES), f"{name}_input")]
        self.input = self.inputs[0]
        self.trainable = True
        self.stop_training = False
        self.optimizer = None
        self._loss_weights = None
        self._lambdas = None

    # ---- shapes -------------------------------------------------------------------------
    def _set_outputs(self, profileNames, countsNames):
        self.outputs = [_TensorSpec((None, self.outputLength, t), n)
                        for t, n in zip(self.headTasks, profileNames)] + \
                       [_TensorSpec((None, 1), n) for n in countsNames]
        self.output = self.outputs
        self.output_names = [o.name for o in self.outputs]

    # ---- Keras-like API -----------------------------------------------------------------
    def compile(self, optimizer=None, loss=None, loss_weights=None, metrics=None):  # noqa: A003
        """``loss`` is accepted for signature compatibility; the losses of this path are fixed:
        multinomial NLL per profile head, lambda-weighted MSE per counts head
        (training.py:22-65).  ``loss_weights`` = profile weights followed by ones."""
        del loss, metrics
        self.optimizer = optimizer if optimizer is not None else _Optimizer()
        H = len(self.headTasks)
        lw = list(loss_weights) if loss_weights is not None else [1.0] * (2 * H)
        self._loss_weights = [float(v) for v in lw[:H]]
        lam = []
        for h in self.headList:
            v = h.get("INTERNAL_λ-variable", h.get("counts-loss-weight", 1.0))
            lam.append(float(v.value if hasattr(v, "value") else v))
        self._lambdas = lam

    def _push_weights(self, net):
        """Loss weights -> device vectors of the engine, only when they changed."""
        key = (tuple(self._loss_weights), tuple(self._lambdas))
        if getattr(self, "_pushed", None) != (id(net), key):
            net.profile_weight.copy_(torch.tensor(self._loss_weights))
            net.lam.copy_(torch.tensor(self._lambdas))
            self._pushed = (id(net), key)

    def set_lambdas(self, lambdas):
        """The adaptive counts-loss weights (callbacks.py:96-389 adjust them between epochs)."""
        self._lambdas = [float(v) for v in lambdas]

    def summary(self, print_fn=print):
        print_fn(f'Model: "{self.name}"')
        for name, shape, n in self._layer_table():
            print_fn(f"  {name:40s} {str(shape):24s} {n:10d}")
        print_fn(f"Total params: {sum(r[2] for r in self._layer_table())}")

    def count_params(self):
        return sum(r[2] for r in self._layer_table())

    def predict(self, x, batch_size=64, verbose=0):
        """(n, inputLength, 4) one-hot -> [profile_h (n, Lout, T_h) ..., counts_h (n, 1) ...]
        as numpy float32 (Keras ``predict`` contract used at utils.py:1021-1023)."""
        del verbose
        dev = self._device
        host = getattr(self, "_predict_host", None)
        if host is not None and not (isinstance(x, torch.Tensor) and x.is_cuda):
            # streaming path: copies overlap the forward passes
            xh = torch.as_tensor(np.ascontiguousarray(x))
            if xh.dtype != torch.uint8:
                xh = xh.to(torch.uint8)
            logits, lc = host(xh, batch_size)
            return self._split_outputs(logits.numpy(), lc.numpy())
        xt = torch.as_tensor(np.ascontiguousarray(x)).to(device=dev, dtype=torch.uint8)
        logits, lc = self._predict_device(xt, batch_size)
        return self._split_outputs(logits.cpu().numpy(), lc.cpu().numpy())

    def _split_outputs(self, logits, lc):
        outs, k = [], 0
        for t in self.headTasks:
            outs.append(np.ascontiguousarray(logits[:, :, k:k + t]))
            k += t
        for h in range(len(self.headTasks)):
            outs.append(np.ascontiguousarray(lc[:, h:h + 1]))
        return outs

    def __call__(self, x, training=False):
        del training
        return self.predict(x[0] if isinstance(x, (list, tuple)) else x)

    # ---- training -----------------------------------------------------------------------
    def fit(self, x, epochs=1, validation_data=None, callbacks=None, verbose=0,
            shuffle=True, initial_epoch=0):
        """Minimal ``Model.fit`` over ``keras.utils.Sequence``-style generators
        (``len(gen)`` batches, ``gen[i] -> (x, (profiles..., logcounts...))``,
        ``gen.on_epoch_end()``), as driven by training.py:168-170.

        Logs per epoch: ``loss`` (weighted total), ``{output}_loss`` per output and their
        ``val_`` twins; callbacks get ``on_train_begin / on_epoch_begin / on_epoch_end /
        on_train_end`` with the Keras signatures and may set ``model.stop_training``."""
        del verbose
        if self.optimizer is None:
            raise RuntimeError("compile() the model before fit()")
        callbacks = list(callbacks or [])
        hist = History()
        for cb in callbacks:
            if hasattr(cb, "set_model"):
                cb.set_model(self)
            if hasattr(cb, "params"):  # Keras' CallbackList.set_params
                cb.params = {"epochs": epochs, "steps": len(x), "verbose": 0}
            if hasattr(cb, "on_train_begin"):
                cb.on_train_begin({})
        batch_cbs = [cb for cb in callbacks if hasattr(cb, "on_train_batch_end")]
        test_cbs = [cb for cb in callbacks if hasattr(cb, "on_test_batch_end")]
        self.stop_training = False
        rng = np.random.default_rng()
        for epoch in range(initial_epoch, epochs):
            for cb in callbacks:
                if hasattr(cb, "on_epoch_begin"):
                    cb.on_epoch_begin(epoch, {})
            order = np.arange(len(x))
            if shuffle:
                rng.shuffle(order)  # Keras shuffles the batch order of a Sequence by default
            logs = self._run_epoch(x, order, train=True, prefix="", batch_cbs=batch_cbs)
            if validation_data is not None:
                logs.update(self._run_epoch(validation_data, np.arange(len(validation_data)),
                                            train=False, prefix="val_", batch_cbs=test_cbs))
            for gen in (x, validation_data):
                if gen is not None and hasattr(gen, "on_epoch_end"):
                    gen.on_epoch_end()
            for cb in callbacks:
                if hasattr(cb, "on_epoch_end"):
                    cb.on_epoch_end(epoch, logs)
            # Keras' History runs after the user callbacks: it records the logs as they left them
            # (FixLossCallback's totals, ReduceLROnPlateau's learning_rate)
            hist.append(epoch, logs)
            if self.stop_training:
                break
        for cb in callbacks:
            if hasattr(cb, "on_train_end"):
                cb.on_train_end({})
        self.history = hist
        return hist

    def _refresh_lambdas(self):
        """Pick up counts-loss weights changed between epochs through the heads'
        ``INTERNAL_λ-variable`` (training.buildLosses / callbacks.ApplyAdaptiveCountsLoss)."""
        if self._lambdas is None:
            return
        for i, h in enumerate(self.headList):
            v = h.get("INTERNAL_λ-variable")
            if v is not None:
                self._lambdas[i] = float(v.read_value() if hasattr(v, "read_value") else v)

    def _run_epoch(self, gen, order, train, prefix, batch_cbs=()):
        H = len(self.headTasks)
        self._refresh_lambdas()
        sums = np.zeros(2 * H, dtype=np.float64)
        sums_dev = None  # losses stay on the device until the epoch ends (one read-back)
        nseq = 0
        for step, bi in enumerate(order):
            xb, yb = gen[int(bi)]
            # log-count targets as the generator supplies them (generators.py:139-140): they are
            # the MSE targets even when they are not log(sum of the profile)
            lct = None
            if isinstance(xb, torch.Tensor):  # device-resident generator: stay on the device
                counts = yb[0] if H == 1 else torch.cat(list(yb[:H]), dim=2)
                if len(yb) >= 2 * H:
                    lct = torch.stack([y.reshape(-1).float() for y in yb[H:2 * H]], dim=1)
            else:
                counts = np.concatenate([np.asarray(y, dtype=np.float32) for y in yb[:H]], axis=2)
                xb = np.asarray(xb)
                if len(yb) >= 2 * H:
                    lct = torch.as_tensor(np.stack(
                        [np.asarray(y, dtype=np.float32).reshape(-1) for y in yb[H:2 * H]], axis=1))
            losses = self._step(xb, counts, train, lct)
            n = xb.shape[0]
            # every loss is a mean over the batch (losses.py:36-38)
            if isinstance(losses, torch.Tensor):
                sums_dev = losses * n if sums_dev is None else sums_dev + losses * n
            else:
                sums += losses * n
            nseq += n
            if batch_cbs:
                # Keras hands the running means to on_train_batch_end / on_test_batch_end; here
                # they are computed only when a callback reads them (a device read-back)
                lazy = _LazyBatchLogs(self, sums, sums_dev, nseq)
                for cb in batch_cbs:
                    (cb.on_train_batch_end if train else cb.on_test_batch_end)(step, lazy)
        if sums_dev is not None:
            sums += sums_dev.cpu().numpy()
        per = sums / max(nseq, 1)
        return self._epoch_logs(per, prefix)

    def _epoch_logs(self, per, prefix):
        H = len(self.headTasks)
        logs = {}
        tot = 0.0
        for h in range(H):
            logs[f"{prefix}{self.output_names[h]}_loss"] = per[h]
            logs[f"{prefix}{self.output_names[H + h]}_loss"] = per[H + h]
            # the metric names Keras derives from metrics=losses, which the reference's callbacks
            # match by regex (callbacks.py:56-59): unweighted MNLL, lambda-weighted MSE
            logs[f"{prefix}{self._metric_base(h, 'profile')}_multinomial_nll"] = per[h]
            logs[f"{prefix}{self._metric_base(h, 'logcounts')}_reweightable_mse"] = per[H + h]
            tot += self._loss_weights[h] * per[h] + per[H + h]
        logs[f"{prefix}loss"] = tot
        return logs

    def _metric_base(self, h, kind):
        """Stem of the per-head metric keys.  Keras derives them from the output layer's name;
        the reference's callbacks find them with ``.*profile_{head}_multinomial_nll`` /
        ``.*logcounts_{head}_reweightable_mse`` under ``re.fullmatch`` (callbacks.py:56-59,
        205-208), which a single-term transformation output (``regress_linear_profile_{head}_r2``,
        layers.py:46) does not satisfy -- the reference raises "The loss didn't match any regex"
        there.  Here the stem always ends in ``{kind}_{head}``, so every model kind trains."""
        name = self.output_names[h if kind == "profile" else len(self.headTasks) + h]
        tail = f"{kind}_{self.headNames[h]}"
        return name if name.endswith(tail) else name[:name.rfind(tail) + len(tail)]

    # ---- weights ------------------------------------------------------------------------
    def save(self, path):
        """``.keras`` (Keras 3 zip: config.json + model.weights.h5), ``.h5`` (legacy whole-model
        HDF5) -- both written without h5py through ``h5lite`` -- or ``.npz`` (native)."""
        arch = self.get_config()
        named = self.named_weights()
        keys = [k for k, _ in named]
        if len(set(keys)) != len(keys):
            dup = sorted({k for k in keys if keys.count(k) > 1})
            raise RuntimeError(f"weight names are not unique: {dup[:4]}")
        if path.endswith(".npz"):
            np.savez(path, __config__=json.dumps(arch), **{k: v for k, v in named})
            return
        from . import keras_io
        keras_io.save_keras(path, arch, named)

    def named_weights(self):
        raise NotImplementedError

    def get_weights(self):
        return [v for _, v in self.named_weights()]


class SoloModel(Model):
    """``soloModel``: one BPNet tower + heads on one GPU."""

    def __init__(self, inputLength, outputLength, numFilters, numLayers, inputFilterWidth,
                 outputFilterWidth, headList, modelName, precise=False, device=None, seed=None):
        super().__init__(f"{modelName}_model", inputLength, outputLength, headList)
        self.modelName = modelName
        self._device = _device(device)
        cfg = NetConfig(inputLength=inputLength, outputLength=outputLength, numFilters=numFilters,
                        numLayers=numLayers, inputFilterWidth=inputFilterWidth,
                        outputFilterWidth=outputFilterWidth, headTasks=self.headTasks,
                        precise=precise)
        self.net = SoloNet(cfg, self._device)  # raises on an inconsistent architecture
        self.net.init_random(seed=seed if seed is not None else int.from_bytes(os.urandom(4), "little"))
        self.inputs = [_TensorSpec((None, inputLength, NUM_BASES), f"{modelName}_input")]
        self.input = self.inputs[0]
        self._set_outputs([f"solo_profile_{n}" for n in self.headNames],
                          [f"solo_logcounts_{n}" for n in self.headNames])

    def get_config(self):
        c = self.net.cfg
        return {"kind": "solo", "modelName": self.modelName, "inputLength": c.inputLength,
                "outputLength": c.outputLength, "numFilters": c.numFilters,
                "numLayers": c.numLayers, "inputFilterWidth": c.inputFilterWidth,
                "outputFilterWidth": c.outputFilterWidth,
                "heads": [{"head-name": n, "num-tasks": t}
                          for n, t in zip(self.headNames, self.headTasks)]}

    def named_weights(self):
        """(Keras variable path, array) in Keras layer order (models.py:85-113)."""
        sd = self.net.state()
        mn = self.modelName
        out = [(f"{mn}_initial_conv/kernel", sd["conv1_w"]), (f"{mn}_initial_conv/bias", sd["conv1_b"])]
        for i in range(self.net.cfg.numLayers):
            out += [(f"{mn}_conv_{i}/kernel", sd[f"dil{i}_w"]), (f"{mn}_conv_{i}/bias", sd[f"dil{i}_b"])]
        for h, n in enumerate(self.headNames):
            out += [(f"solo_profile_{n}/kernel", sd[f"prof{h}_w"]),
                    (f"solo_profile_{n}/bias", sd[f"prof{h}_b"])]
            out += [(f"solo_logcounts_{n}/kernel", sd[f"cnt{h}_w"]),
                    (f"solo_logcounts_{n}/bias", sd[f"cnt{h}_b"])]
        return out

    def set_weights(self, weights):
        names = [k for k, _ in self.named_weights()]
        if isinstance(weights, dict):
            weights = [weights[k] for k in names]
        assert len(weights) == len(names), "weight list does not match the architecture"
        sd = {}
        it = iter(weights)
        sd["conv1_w"], sd["conv1_b"] = next(it), next(it)
        for i in range(self.net.cfg.numLayers):
            sd[f"dil{i}_w"], sd[f"dil{i}_b"] = next(it), next(it)
        for h in range(len(self.headNames)):
            sd[f"prof{h}_w"], sd[f"prof{h}_b"] = next(it), next(it)
            sd[f"cnt{h}_w"], sd[f"cnt{h}_b"] = next(it), next(it)
        self.net.load_state(sd)

    def _layer_table(self):
        return [(k, v.shape, int(v.size)) for k, v in self.named_weights()]

    def _predict_device(self, xt, batch_size):
        return self.net.predict(xt, batchSize=batch_size)

    def _predict_host(self, xh, batch_size):
        return self.net.predict_host(xh, batchSize=batch_size)

    def _step(self, xb, counts, train, lct=None):
        net = self.net
        net.enable_training()
        self._push_weights(net)
        if train and self.trainable:
            losses = net.train_step(torch.as_tensor(xb), torch.as_tensor(counts),
                                    self.optimizer.learning_rate, logcounts_true=lct)
        else:
            losses = net.eval_losses(torch.as_tensor(xb), torch.as_tensor(counts),
                                     logcounts_true=lct)
        return losses.double()  # a copy: the engine reuses its loss buffer next step


# ----------------------------------------------------------------------------------------
# transformation model (models.py:118-265)
# ----------------------------------------------------------------------------------------
_ACT = {"linear": 0, "sigmoid": 1, "relu": 2}


class TransformationModel(Model):
    """``transformationModel``: per head, sums of ``m1 * act(m2 * u + b2) + b1`` applied to the
    frozen solo model's profile logits and log-counts.  Parameters are scalars initialised to
    slope 1 / offset 0 (layers.py:37-46 ``kernel_initializer="Ones"``, zero bias)."""

    def __init__(self, soloModelIn, profileArchitectureSpecification,
                 countsArchitectureSpecification, headList):
        super().__init__("transformation_model", soloModelIn.inputLength,
                         soloModelIn.outputLength, headList)
        soloModelIn.trainable = False  # models.py:249
        self.solo = soloModelIn
        self._device = soloModelIn._device
        self.profileSpec = profileArchitectureSpecification
        self.countsSpec = countsArchitectureSpecification
        for spec in (self.profileSpec, self.countsSpec):
            if spec["name"] not in ("simple", "passthrough"):
                raise ValueError("Currently, only simple regression is supported.")
            for t in spec.get("types", []):
                if t not in _ACT:
                    raise ValueError(f"The simple layer type you gave ({t}) is not supported")
        H = len(self.headTasks)
        dev = self._device
        # coefficient table [(head, kind) -> terms][4] = {m1, b1, m2, b2}
        self.pTypes = list(self.profileSpec.get("types", [])) if self.profileSpec["name"] == "simple" else []
        self.cTypes = list(self.countsSpec.get("types", [])) if self.countsSpec["name"] == "simple" else []
        n_terms = H * (len(self.pTypes) + len(self.cTypes))
        base = torch.tensor([1.0, 0.0, 1.0, 0.0], dtype=torch.float32)
        self.coeffs = base.repeat(max(n_terms, 1), 1).to(dev).contiguous()
        self.pAct = torch.tensor([_ACT[t] for t in self.pTypes] or [0], dtype=torch.int32, device=dev)
        self.cAct = torch.tensor([_ACT[t] for t in self.cTypes] or [0], dtype=torch.int32, device=dev)
        self._adam = None
        self._set_outputs([self._out_name("profile", n, self.pTypes, f"solo_profile_{n}")
                           for n in self.headNames],
                          [self._out_name("logcounts", n, self.cTypes, f"solo_logcounts_{n}")
                           for n in self.headNames])

    @staticmethod
    def _out_name(kind, head, types, passthrough):
        if not types:
            return passthrough
        if len(types) > 1:
            return f"regress_sum_{kind}_{head}"
        t = types[0]
        return (f"regress_linear_{kind}_{head}_r2" if t == "linear"
                else f"{t}_out_linear_{kind}_{head}_r2")

    def _rows(self, h, kind):
        """Slice of the coefficient table holding head h's profile / counts terms."""
        per = len(self.pTypes) + len(self.cTypes)
        off = h * per + (0 if kind == "p" else len(self.pTypes))
        return off, off + (len(self.pTypes) if kind == "p" else len(self.cTypes))

    def get_config(self):
        return {"kind": "transformation", "solo": self.solo.get_config(),
                "profile": self.profileSpec, "counts": self.countsSpec}

    def named_weights(self):
        out = list(self.solo.named_weights())
        c = self.coeffs.cpu().numpy()
        for h, n in enumerate(self.headNames):
            for kind, types, tag in (("p", self.pTypes, "profile"), ("c", self.cTypes, "logcounts")):
                lo, _ = self._rows(h, kind)
                for i, t in enumerate(types):
                    m1, b1, m2, b2 = c[lo + i]
                    if t == "linear":
                        out += [(f"regress_linear_{tag}_{n}_conv/kernel", np.float32(m1).reshape(1, 1, 1)),
                                (f"regress_linear_{tag}_{n}_conv/bias", np.float32(b1).reshape(1))]
                    else:
                        out += [(f"{t}_in_linear_{tag}_{n}_conv/kernel", np.float32(m2).reshape(1, 1, 1)),
                                (f"{t}_in_linear_{tag}_{n}_conv/bias", np.float32(b2).reshape(1)),
                                (f"{t}_out_linear_{tag}_{n}_conv/kernel", np.float32(m1).reshape(1, 1, 1)),
                                (f"{t}_out_linear_{tag}_{n}_conv/bias", np.float32(b1).reshape(1))]
        return out

    def set_weights(self, weights):
        names = [k for k, _ in self.named_weights()]
        if isinstance(weights, dict):
            weights = [weights[k] for k in names]
        n_solo = len(self.solo.named_weights())
        self.solo.set_weights(list(weights[:n_solo]))
        vals = dict(zip(names[n_solo:], weights[n_solo:]))
        c = self.coeffs.cpu().numpy().copy()
        for h, n in enumerate(self.headNames):
            for kind, types, tag in (("p", self.pTypes, "profile"), ("c", self.cTypes, "logcounts")):
                lo, _ = self._rows(h, kind)
                for i, t in enumerate(types):
                    if t == "linear":
                        c[lo + i, 0] = np.ravel(vals[f"regress_linear_{tag}_{n}_conv/kernel"])[0]
                        c[lo + i, 1] = np.ravel(vals[f"regress_linear_{tag}_{n}_conv/bias"])[0]
                    else:
                        c[lo + i, 2] = np.ravel(vals[f"{t}_in_linear_{tag}_{n}_conv/kernel"])[0]
                        c[lo + i, 3] = np.ravel(vals[f"{t}_in_linear_{tag}_{n}_conv/bias"])[0]
                        c[lo + i, 0] = np.ravel(vals[f"{t}_out_linear_{tag}_{n}_conv/kernel"])[0]
                        c[lo + i, 1] = np.ravel(vals[f"{t}_out_linear_{tag}_{n}_conv/bias"])[0]
        self.coeffs.copy_(torch.tensor(c))

    def _layer_table(self):
        return [(k, np.shape(v), int(np.size(v))) for k, v in self.named_weights()]

    # ---- forward on device tensors ------------------------------------------------------
    def transform_device(self, logits, lc):
        """Apply the per-head transformations to solo outputs (B, L, T) / (B, H) on the device.
        Returns new tensors (the inputs are kept for the backward pass)."""
        out_l = logits.clone()
        out_c = lc.clone()
        k = 0
        for h, t in enumerate(self.headTasks):
            if self.pTypes:
                lo, hi = self._rows(h, "p")
                u = logits[:, :, k:k + t].contiguous()
                y = torch.empty_like(u)
                ops.transform_fwd(u, y, self.pAct, self.coeffs[lo:hi].contiguous())
                out_l[:, :, k:k + t] = y
            if self.cTypes:
                lo, hi = self._rows(h, "c")
                u = lc[:, h:h + 1].contiguous()
                y = torch.empty_like(u)
                ops.transform_fwd(u, y, self.cAct, self.coeffs[lo:hi].contiguous())
                out_c[:, h:h + 1] = y
            k += t
        return out_l, out_c

    def _predict_device(self, xt, batch_size):
        logits, lc = self.solo.net.predict(xt, batchSize=batch_size)
        return self.transform_device(logits, lc)

    def _step(self, xb, counts, train, lct=None):
        """One batch of trainTransformationModel: solo forward (frozen), transformations, losses,
        gradients of the 4-vectors, Keras-Adam on them."""
        net = self.solo.net
        net.enable_training()
        dev = self._device
        xt = torch.as_tensor(xb).to(device=dev, dtype=torch.uint8)
        ct = torch.as_tensor(counts).to(device=dev, dtype=torch.float32).contiguous()
        B = xt.shape[0]
        H = len(self.headTasks)
        logits, lc = net.predict(xt, batchSize=xt.shape[0])
        tl, tc = self.transform_device(logits, lc)
        losses = torch.zeros((2 * H,), dtype=torch.float32, device=dev)
        dl = torch.empty_like(tl)
        dc = torch.empty_like(tc)
        pw = torch.tensor(self._loss_weights, dtype=torch.float32, device=dev)
        lam = torch.tensor(self._lambdas, dtype=torch.float32, device=dev)
        if lct is not None:
            lct = torch.as_tensor(lct).to(device=dev, dtype=torch.float32).reshape(B, H).contiguous()
        ops.loss_fwd_bwd(net.task_off, tl.contiguous(), tc.contiguous(), ct, pw, lam, losses, dl, dc,
                         logcounts_true=lct)
        if train:
            grads = torch.zeros_like(self.coeffs)
            k = 0
            for h, t in enumerate(self.headTasks):
                if self.pTypes:
                    lo, hi = self._rows(h, "p")
                    g = torch.zeros((hi - lo, 4), dtype=torch.float32, device=dev)
                    ops.transform_bwd(logits[:, :, k:k + t].contiguous(), dl[:, :, k:k + t].contiguous(),
                                      self.pAct, self.coeffs[lo:hi].contiguous(), g)
                    grads[lo:hi] = g
                if self.cTypes:
                    lo, hi = self._rows(h, "c")
                    g = torch.zeros((hi - lo, 4), dtype=torch.float32, device=dev)
                    ops.transform_bwd(lc[:, h:h + 1].contiguous(), dc[:, h:h + 1].contiguous(),
                                      self.cAct, self.coeffs[lo:hi].contiguous(), g)
                    grads[lo:hi] = g
                k += t
            # a linear term has no inner regression: its (m2, b2) slots are not parameters
            for h in range(H):
                for kind, types in (("p", self.pTypes), ("c", self.cTypes)):
                    lo, _ = self._rows(h, kind)
                    for i, ty in enumerate(types):
                        if ty == "linear":
                            grads[lo + i, 2:] = 0.0
            if self._adam is None:
                self._adam = {"m": torch.zeros_like(self.coeffs), "v": torch.zeros_like(self.coeffs),
                              "sl": torch.zeros((2,), dtype=torch.float32, device=dev)}
            a = self._adam
            a["sl"][0] += 1.0
            a["sl"][1] = float(self.optimizer.learning_rate)
            flat = self.coeffs.view(-1)
            ops.adam_step(flat, grads.view(-1).contiguous(), a["m"].view(-1), a["v"].view(-1), a["sl"])
        return losses.double().cpu().numpy()


# ----------------------------------------------------------------------------------------
# combined model (models.py:268-387)
# ----------------------------------------------------------------------------------------
class CombinedModel(Model):
    """``combinedModel``: frozen (solo + transformation) bias model on the cropped input, plus a
    trainable residual BPNet; logits add, log-counts combine as log(e^bias + e^resid) in fp64
    where ``use-bias-counts`` (layers.py:52-75, models.py:363-380)."""

    def __init__(self, residual, biasModel, headList):
        super().__init__("combined_model", residual.inputLength, residual.outputLength, headList)
        biasModel.trainable = False
        self.residual = residual
        self.bias = biasModel
        self._device = residual._device
        cropDiff = residual.inputLength - biasModel.inputLength
        assert cropDiff % 2 == 0, "Cropping returns an odd number: " \
                                  "redo your model input sizes to be even numbers."
        assert cropDiff >= 0, "Bias model inputs are larger than residual inputs"
        self.cropFlank = cropDiff // 2
        assert biasModel.outputLength == residual.outputLength, \
            "bias and residual models must predict the same output length"
        self.useBias = torch.tensor([1 if h.get("use-bias-counts", False) else 0 for h in headList],
                                    dtype=torch.int32, device=self._device)
        self.inputs = residual.inputs
        self.input = residual.input
        self._set_outputs([f"combined_add_profile_{n}" for n in self.headNames],
                          [f"combined_logcounts_{n}" for n in self.headNames])

    def get_config(self):
        return {"kind": "combined", "residual": self.residual.get_config(),
                "bias": self.bias.get_config(),
                "use-bias-counts": [bool(v) for v in self.useBias.cpu().tolist()]}

    def named_weights(self):
        """Residual variables under their own layer names, the frozen bias model's under the
        nested model's name (in the reference graph it IS a nested ``transformation_model``,
        models.py:349-352): both towers own layers called ``solo_profile_{head}`` /
        ``solo_logcounts_{head}`` (models.py:41-48), so a flat list would collide."""
        return list(self.residual.named_weights()) + \
            [(f"{self.bias.name}/{k}", v) for k, v in self.bias.named_weights()]

    def set_weights(self, weights):
        n = len(self.residual.named_weights())
        self.residual.set_weights(list(weights[:n]))
        self.bias.set_weights(list(weights[n:]))

    def _layer_table(self):
        return [(k, np.shape(v), int(np.size(v))) for k, v in self.named_weights()]

    def _bias_parts(self):
        """(solo network of the bias model, profile act table, counts act table, coefficient
        table, n profile terms, n counts terms) -- a plain solo bias model has no terms."""
        if isinstance(self.bias, TransformationModel):
            tm = self.bias
            return (tm.solo.net, tm.pAct, tm.cAct, tm.coeffs, len(tm.pTypes), len(tm.cTypes))
        return (self.bias.net, None, None, None, 0, 0)

    def _bias_hook(self, B, kind):
        """The closure ``SoloNet.forward_backward_into`` / ``eval_losses`` call after the residual
        forward: crop the input (models.py:339-347), run the frozen bias network, then ONE launch
        for transformation + logit add + CountsLogSumExp (``bpb_combine_bias``) into the
        residual workspace.  Every buffer is fixed, so the whole combined step -- bias forward
        included -- replays as one CUDA graph, and shards over GPUs like the solo step."""
        key = (B, kind)
        hooks = self.__dict__.setdefault("_hooks", {})
        if key not in hooks:
            bnet, pAct, cAct, coeffs, n_p, n_c = self._bias_parts()
            wb = bnet._workspace(B, "predict")
            f = self.cropFlank
            useBias = self.useBias

            def hook(ws):
                x = ws["x"]
                wb["x"].copy_(x[:, f:x.shape[1] - f, :] if f > 0 else x)
                bnet.predict_into(wb)
                ops.combine_bias(bnet.task_off, wb["logits"], wb["logcounts"], pAct, cAct, coeffs,
                                 n_p, n_c, useBias, ws["logits"], ws["logcounts"],
                                 ws["logits_c"], ws["lc_c"], ws["dscale"])
            hooks[key] = hook
        return hooks[key]

    def _predict_device(self, xt, batch_size):
        bnet, pAct, cAct, coeffs, n_p, n_c = self._bias_parts()
        f = self.cropFlank
        xc = xt[:, f:xt.shape[1] - f, :].contiguous() if f > 0 else xt
        bl, bc = bnet.predict(xc, batchSize=batch_size)
        rl, rc = self.residual.net.predict(xt, batchSize=batch_size)
        out_l, out_c = torch.empty_like(rl), torch.empty_like(rc)
        if rl.shape[0] > 0:
            ops.combine_bias(bnet.task_off, bl, bc, pAct, cAct, coeffs, n_p, n_c, self.useBias,
                             rl, rc, out_l, out_c)
        return out_l, out_c

    def _step(self, xb, counts, train, lct=None, peers=None, allreduce=None, world=1):
        net = self.residual.net
        net.enable_training()
        self._push_weights(net)
        xt, ct = torch.as_tensor(xb), torch.as_tensor(counts)
        if xt.dtype != torch.uint8:
            xt = xt.to(torch.uint8)
        if lct is not None:
            lct = torch.as_tensor(lct).to(device=self._device, dtype=torch.float32)
        B = xt.shape[0]
        if train:
            losses = net.train_step(xt, ct.float(), self.optimizer.learning_rate,
                                    bias_logits=self._bias_hook(B, "train"), logcounts_true=lct,
                                    peers=peers, allreduce=allreduce, world=world)
        else:
            losses = net.eval_losses(xt.to(self._device), ct.to(self._device, torch.float32),
                                     bias_logits=self._bias_hook(B, "eval"), logcounts_true=lct)
        return losses.double()


# ----------------------------------------------------------------------------------------
# the reference's builder functions
# ----------------------------------------------------------------------------------------
def soloModel(inputLength, outputLength, numFilters, numLayers, inputFilterWidth,
              outputFilterWidth, headList, modelName, **kw):
    """models.py:53-115.  It is an error to call this with an inconsistent network structure
    (the reference lets Keras raise; here ``NetConfig.validate`` raises ``ValueError``)."""
    return SoloModel(inputLength, outputLength, numFilters, numLayers, inputFilterWidth,
                     outputFilterWidth, headList, modelName, **kw)


def transformationModel(soloModelIn, profileArchitectureSpecification,
                        countsArchitectureSpecification, headList):
    """models.py:226-265."""
    return TransformationModel(soloModelIn, profileArchitectureSpecification,
                               countsArchitectureSpecification, headList)


def combinedModel(inputLength, outputLength, numFilters, numLayers, inputFilterWidth,
                  outputFilterWidth, headList, biasModel, **kw):
    """models.py:268-387.  Returns (combined, residual, biasHeads) like the reference; the third
    element is the (frozen) bias model applied to the cropped input."""
    kw.setdefault("device", biasModel._device)
    residual = soloModel(inputLength, outputLength, numFilters, numLayers, inputFilterWidth,
                         outputFilterWidth, headList, "residual", **kw)
    combined = CombinedModel(residual, biasModel, headList)
    return combined, residual, biasModel


def loadModel(path, device=None, precise=False):
    """Counterpart of ``utils.loadModel`` (utils.py:24-85): ``.keras`` and legacy ``.h5`` files
    written by Keras for the reference's builders or by ``Model.save`` here (the architecture is
    derived from the file's own layer configuration), and the native ``.npz``."""
    if path.endswith(".model") and not os.path.exists(path):
        path = path[:-5] + "keras"  # utils.py:67-72: old extension in an old configuration file
    if path.endswith(".npz"):
        z = np.load(path, allow_pickle=False)
        arch = json.loads(str(z["__config__"]))
        weights = {k: z[k] for k in z.files if k != "__config__"}
    else:
        from . import keras_io
        arch, weights = keras_io.load_keras(path)
    model = _build_from_config(arch, device, precise)
    model.set_weights([weights[k] for k, _ in model.named_weights()])
    return model


def _build_from_config(arch, device, precise):
    if arch["kind"] == "solo":
        return soloModel(arch["inputLength"], arch["outputLength"], arch["numFilters"],
                         arch["numLayers"], arch["inputFilterWidth"], arch["outputFilterWidth"],
                         arch["heads"], arch["modelName"], device=device, precise=precise, seed=0)
    if arch["kind"] == "transformation":
        solo = _build_from_config(arch["solo"], device, precise)
        return transformationModel(solo, arch["profile"], arch["counts"], arch["solo"]["heads"])
    if arch["kind"] == "combined":
        bias = _build_from_config(arch["bias"], device, precise)
        r = arch["residual"]
        heads = [dict(h, **{"use-bias-counts": u}) for h, u in zip(r["heads"], arch["use-bias-counts"])]
        combined, _, _ = combinedModel(r["inputLength"], r["outputLength"], r["numFilters"],
                                       r["numLayers"], r["inputFilterWidth"],
                                       r["outputFilterWidth"], heads, bias, precise=precise, seed=0)
        return combined
    raise ValueError(f"unknown model kind {arch['kind']!r}")


def lengthDifference(numLayers, inputFilterWidth, outputFilterWidth):
    """lengthCalc.getLengthDifference (lengthCalc.py:71-123) for a solo model."""
    return (inputFilterWidth - 1) + sum(2 * 2 ** (i + 1) for i in range(numLayers)) + \
        (outputFilterWidth - 1)
