"""Drop-in surface of the reference's ``models/wgancls/model.py`` (class WGanCls, :5-225) on the
B200-native engine.  Same constructor, same attribute names for inputs ("placeholders"), fetches
and sizes, same ``generator`` / ``discriminator`` signatures; ``run(fetches, feed_dict)`` stands in
for ``sess.run`` with exactly the trainer's fetch lists (models/wgancls/trainer.py:97,101,111).

All arithmetic happens in the CUDA library (text-to-image_b200/csrc) reached through
``t2i_b200.kernels``; PyTorch provides device memory, streams and (multi-GPU) torch.distributed.
There is no CPU fallback: constructing the training graph without the library or a GPU raises.
"""
import math

import numpy as np
import torch

from ...engine import Engine, KT_INIT

NHWC = "NHWC"
NCHW = "NCHW"


class Placeholder(object):
    """Feed key standing in for a tf.placeholder (models/wgancls/model.py:36-46)."""

    def __init__(self, name, shape=None):
        self.name, self.shape = name, shape

    def __repr__(self):
        return "<placeholder %s %s>" % (self.name, self.shape)


class Fetch(object):
    """Fetchable graph element: a train op (returns None) or a named scalar / tensor."""

    def __init__(self, name, kind):
        self.name, self.kind = name, kind

    def __repr__(self):
        return "<fetch %s>" % self.name


def _truncated_normal(shape, device, generator=None):
    """tf.truncated_normal (model.py:119): N(0,1) re-drawn outside +-2 sigma."""
    t = torch.empty(shape, device=device, dtype=torch.float32)
    torch.nn.init.trunc_normal_(t, 0.0, 1.0, -2.0, 2.0, generator=generator)
    return t


class WGanCls(object):
    def __init__(self, cfg, build_model=True, precision="bf16", device=None, kernels=None, distributed=None,
                 use_graphs=True, sync_bn=False):
        """
        Args:
          cfg: Config specifying all the parameters of the model (reference: model.py:6-10).
          build_model: allocate the training graph (engine buffers for TRAIN.BATCH_SIZE); with False
            only parameters exist and ``generator``/``discriminator`` allocate per call (eval / vis).
          precision: "bf16" (throughput: bf16 storage, fp32 accumulate) or "bf16x3" (parity: split
            bf16, three tensor-core products per contraction, ~fp32-faithful).
          distributed: None, or a torch.distributed process group handle/True for batch sharding with
            one gradient allreduce per optimizer step.
          use_graphs: capture the D run and the G run into CUDA graphs after their first eager call.
          sync_bn: with ``distributed``, all-reduce g_net's BatchNorm sums so that the statistics are those of
            the GLOBAL batch, exactly what the single-device reference computes (utils/ops.py:20-29); default
            False = per-replica statistics and one all-reduce per optimizer step.
        """
        self.cfg = cfg

        self.batch_size = cfg.TRAIN.BATCH_SIZE
        self.sample_num = cfg.TRAIN.SAMPLE_NUM

        self.output_size = cfg.MODEL.OUTPUT_SIZE

        self.z_dim = cfg.MODEL.Z_DIM
        self.embed_dim = cfg.MODEL.EMBED_DIM
        self.compressed_embed_dim = cfg.MODEL.COMPRESSED_EMBED_DIM

        self.gf_dim = cfg.MODEL.GF_DIM
        self.df_dim = cfg.MODEL.DF_DIM

        self.image_dims = [cfg.MODEL.IMAGE_SHAPE.H, cfg.MODEL.IMAGE_SHAPE.W, cfg.MODEL.IMAGE_SHAPE.D]
        if self.output_size != 64 or self.image_dims != [64, 64, 3]:
            # the reference hard-codes the 4x4 embedding tile (model.py:154): only 64x64x3 is valid
            raise ValueError("wgancls is defined for 64x64x3 images only (got %s)" % (self.image_dims,))

        self.global_step = 0
        self.precision = precision
        self._np = {"bf16": 1, "bf16x3": 2}[precision]
        if kernels is None:
            from ... import kernels as _k      # loads libt2i_b200.so lazily, raises if missing
            kernels = _k
        self._K = kernels
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("WGanCls needs a CUDA device (sm_100a); there is no CPU fallback")
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self._world, self._allreduce = 1, None
        if distributed:
            import torch.distributed as dist
            group = None if distributed is True else distributed
            self._world = dist.get_world_size(group)
            self._allreduce = lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        self._engines = {}
        self._use_graphs = use_graphs
        self._sync_bn = sync_bn
        self._noise_gen = None
        self._built = False

        # inputs (model.py:36-46)
        b, dims = self.batch_size, self.image_dims
        self.iter = Placeholder("iter")
        self.learning_rate_d = Placeholder("learning_rate_d")
        self.learning_rate_g = Placeholder("learning_rate_g")
        self.x = Placeholder("real_images", [b] + dims)
        self.x_mismatch = Placeholder("wrong_images", [b] + dims)
        self.cond = Placeholder("cond", [b, self.embed_dim])
        self.z = Placeholder("z", [b, self.z_dim])
        self.epsilon = Placeholder("eps", [b, 1, 1, 1])
        self.z_sample = Placeholder("z_sample", [self.sample_num, self.z_dim])
        self.cond_sample = Placeholder("cond_sample", [self.sample_num, self.embed_dim])
        # explicit noise inputs (the reference draws them inside the graph, model.py:119); optional
        self.cond_noise = Placeholder("cond_noise", [b, self.compressed_embed_dim])
        self.cond_noise_sample = Placeholder("cond_noise_sample", [self.sample_num, self.compressed_embed_dim])

        self._train_engine()      # parameters exist from construction on, like TF variables
        if build_model:
            self.build_model()
            self.define_losses()

    # ------------------------------------------------------------------ graph
    def _engine(self, batch):
        if batch not in self._engines:
            base = next(iter(self._engines.values()), None)
            self._engines[batch] = Engine(
                self._K, self.device, batch, self._np, self.z_dim, self.embed_dim, self.compressed_embed_dim,
                self.gf_dim, self.df_dim, self.cfg.TRAIN.BETA1, self.cfg.TRAIN.BETA2, self.cfg.TRAIN.COEFF.KL,
                self._world, self._allreduce, share_from=base, use_graphs=self._use_graphs, sync_bn=self._sync_bn)
        return self._engines[batch]

    def _train_engine(self):
        return self._engine(self.batch_size)

    def build_model(self):
        """model.py:34-60.  The fetchable tensors of the reference graph, by name."""
        for n in ("G", "embed_mean", "embed_log_sigma", "Dg_logit", "Dx_logit", "Dxmi_logit", "x_hat",
                  "Dx_hat_logit", "sampler"):
            setattr(self, n, Fetch(n, "tensor"))
        self.cond_inp = self.cond
        self.d_vars = [n for n in self.variable_names() if n.startswith("d_net/")]
        self.g_vars = [n for n in self.variable_names() if n.startswith("g_net/") and "moving_" not in n]
        self._built = True

    def define_losses(self):
        """model.py:72-106.  Loss scalars and the three train ops, by name."""
        self.kt = Fetch("kt", "scalar")
        for n in ("D_loss_real", "D_loss_fake", "D_loss_mismatch", "wdist", "wdist2", "reg_loss", "balance_loss",
                  "G_kl_loss", "real_gp", "real_gp2", "D_loss", "G_loss"):
            setattr(self, n, Fetch(n, "scalar"))
        self.D_optim = Fetch("D_optim", "op")
        self.kt_optim = Fetch("kt_optim", "op")
        self.G_optim = Fetch("G_optim", "op")

    # ------------------------------------------------------------------ variables (checkpoint boundary)
    def variable_names(self):
        return list(self._train_engine().get_params_tf().keys())

    def get_variables(self):
        """All variables in the reference's TF layout and default scope names (+ 'kt', 'global_step')."""
        eng = self._train_engine()
        out = eng.get_params_tf()
        out["kt"] = eng.kt.detach().cpu().reshape(())
        out["global_step"] = torch.tensor(self.global_step)
        return out

    def set_variables(self, variables):
        eng = self._train_engine()
        p = {k: v for k, v in variables.items() if k not in ("kt", "global_step")}
        eng.set_params_tf(p)
        if "kt" in variables:
            eng.kt.fill_(float(variables["kt"]))
        if "global_step" in variables:
            self.global_step = int(variables["global_step"])

    def get_optimizer_state(self):
        """Adam slots (TF layout, 'm/<variable>' / 'v/<variable>') and step counters: what a full
        tf.train.Saver checkpoint holds beside the variables (trainer.py:52)."""
        return self._train_engine().get_adam_tf()

    def set_optimizer_state(self, state):
        if state:
            self._train_engine().set_adam_tf(state)

    def initialize(self, seed=0):
        """tf.global_variables_initializer(): the reference initialisation (utils/ops.py:60,68,86:
        He truncated normal; zero biases; gamma 1, beta 0; moving 0/1; kt 0.7 at model.py:77)."""
        gen = torch.Generator().manual_seed(seed)
        eng = self._train_engine()
        p = eng.get_params_tf()
        for name, w in p.items():
            leaf = name.rsplit("/", 1)[1]
            if leaf in ("weights", "kernel"):
                fan_in = w.shape[-2] * int(np.prod(w.shape[:-2])) if w.dim() > 1 else 1
                std = math.sqrt(1.3 * 2.0 / fan_in)
                t = torch.empty(w.shape)
                torch.nn.init.trunc_normal_(t, 0.0, 1.0, -2.0, 2.0, generator=gen)
                p[name] = t * std
            elif leaf in ("gamma", "moving_variance"):
                p[name] = torch.ones_like(w)
            else:
                p[name] = torch.zeros_like(w)
        eng.set_params_tf(p)
        eng.kt.fill_(KT_INIT)
        eng.adam_m["d"].zero_(); eng.adam_v["d"].zero_(); eng.adam_m["g"].zero_(); eng.adam_v["g"].zero_()
        eng.d_t = eng.g_t = 0
        self.global_step = 0

    # ------------------------------------------------------------------ eager sub-graphs
    def _dev(self, a, shape=None):
        t = torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a)
        t = t.to(self.device, torch.float32, non_blocking=True)
        return t.reshape(shape) if shape is not None else t

    def seed_noise(self, seed):
        """Seed the device generator used for the in-graph truncated-normal draw (model.py:119)."""
        self._noise_gen = torch.Generator(device=self.device).manual_seed(seed)

    def generator(self, z, embed, reuse=False, is_training=True, df=NCHW, cond_noise=True, noise=None):
        """model.py:163-225.  Returns (image NHWC in (-1,1), mean, log_sigma) as CUDA fp32 tensors.
        ``is_training`` selects batch vs moving BatchNorm statistics; ``noise`` optionally supplies
        the truncated-normal draw.  ``reuse``/``df`` are accepted for signature compatibility."""
        z = self._dev(z)
        b = z.shape[0]
        eng = self._engine(b)
        embed = self._dev(embed, (b, self.embed_dim))
        tn = self._dev(noise) if noise is not None else _truncated_normal((b, self.compressed_embed_dim),
                                                                         self.device, self._noise_gen)
        out = torch.empty(b, 64, 64, 3, device=self.device, dtype=torch.float32)
        eng.g["kl_scratch"].zero_()
        eng.g_forward(z, embed, tn, out, eng.g["kl_scratch"], train=is_training, cond_noise=cond_noise)
        ms = eng.g["ms"].clone()        # fp32 [b, 2*ce] = [mean | log_sigma]
        ce = self.compressed_embed_dim
        return out, ms[:, :ce], ms[:, ce:]

    def discriminator(self, inputs, embed, reuse=False):
        """model.py:129-161.  inputs NHWC [B,64,64,3], embed [B,1024] -> logits [B,1,1,1]."""
        x = self._dev(inputs)
        b = x.shape[0]
        eng = self._engine(b)
        eng.d["img"][:b].copy_(x)
        self._K.to_planes(self._dev(embed, (b, self.embed_dim)), eng.d["cond"][:, :b])
        eng.d_forward(0, b)
        return eng.d["logit"][:b].clone().reshape(b, 1, 1, 1)

    # ------------------------------------------------------------------ sess.run stand-in
    def _feed(self, feed_dict, key, default=None):
        return feed_dict.get(key, default)

    def run(self, fetches, feed_dict=None):
        """``sess.run(fetches, feed_dict)`` for the reference trainer's fetch lists:
        [D_optim, kt_optim, D_loss] (trainer.py:97), [G_optim, G_loss] (:101), sampler (:111) and
        the summary scalars (:22-44).  Train ops return None, scalars Python floats, tensors numpy."""
        if not self._built:
            raise RuntimeError("the model was constructed with build_model=False")
        feed_dict = feed_dict or {}
        single = not isinstance(fetches, (list, tuple))
        flist = [fetches] if single else list(fetches)
        names = [f.name for f in flist]
        eng = self._train_engine()
        ran = False
        if "D_optim" in names or "kt_optim" in names:
            self._stage(eng, feed_dict, need_images=True)
            eng.d_step(float(self._feed(feed_dict, self.learning_rate_d, self.cfg.TRAIN.D_LR)))
            self.global_step += 1                                   # model.py:97
            ran = True
        if "G_optim" in names:
            self._stage(eng, feed_dict, need_images=False)
            eng.g_step(float(self._feed(feed_dict, self.learning_rate_g, self.cfg.TRAIN.G_LR)))
            ran = True
        out = []
        sc = None
        for f in flist:
            if f.kind == "op":
                out.append(None)
            elif f.kind == "scalar":
                if not ran and sc is None and f.name != "kt":
                    raise RuntimeError("scalar '%s' is produced by the D/G run; fetch it together with the "
                                       "train op or after it" % f.name)
                if f.name == "kt" and not ran:
                    eng.join_comm()                  # the kt step of the last D run lives on the communication stream
                    out.append(float(eng.kt.item()))
                    continue
                sc = sc or eng.scalars_dict()        # 'kt' there is the value after this iteration's kt_optim
                out.append(sc[f.name])
            elif f.name == "sampler":
                z = self._dev(feed_dict[self.z_sample])
                cond = self._dev(feed_dict[self.cond_sample])
                img, _, _ = self.generator(z, cond, is_training=False,
                                           noise=self._feed(feed_dict, self.cond_noise_sample))
                out.append(img.cpu().numpy())
            elif f.name == "G":
                out.append(eng.d["img"][:self.batch_size].cpu().numpy())
            # the remaining tensors of build_model (model.py:48-55) live in the engine's buffers after a D run:
            # image segments [fake | real | mismatch | x_hat] and one logit per sample of the 4B batch
            elif f.name == "x_hat":
                b = self.batch_size
                out.append(eng.d["img"][3 * b:].cpu().numpy())
            elif f.name in ("Dg_logit", "Dx_logit", "Dxmi_logit", "Dx_hat_logit"):
                b = self.batch_size
                k = ("Dg_logit", "Dx_logit", "Dxmi_logit", "Dx_hat_logit").index(f.name)
                out.append(eng.d["logit"][k * b:(k + 1) * b].cpu().numpy().reshape(b, 1, 1, 1))
            elif f.name in ("embed_mean", "embed_log_sigma"):
                ce = self.compressed_embed_dim
                ms = eng.g["ms"]
                out.append((ms[:, :ce] if f.name == "embed_mean" else ms[:, ce:]).cpu().numpy())
            else:
                raise KeyError("fetch '%s' is not materialised by this implementation" % f.name)
        return out[0] if single else out

    def _stage(self, eng, feed_dict, need_images):
        b = self.batch_size
        get = lambda k: feed_dict[k] if k in feed_dict else None
        noise = get(self.cond_noise)
        if noise is None:                                           # a fresh draw per run
            noise = _truncated_normal((b, self.compressed_embed_dim), self.device, self._noise_gen)
        t = lambda a: None if a is None else torch.as_tensor(np.asarray(a, dtype=np.float32) if not torch.is_tensor(a) else a)
        eng.load_feed(x=t(get(self.x)) if need_images else None,
                      x_mismatch=t(get(self.x_mismatch)) if need_images else None,
                      cond=t(get(self.cond)), z=t(get(self.z)),
                      epsilon=t(get(self.epsilon)) if need_images else None, tn_eps=t(noise))

    # convenience names for direct use (bench.py, tests)
    def d_step(self, feed_dict):
        return self.run([self.D_optim, self.kt_optim, self.D_loss], feed_dict)[2]

    def g_step(self, feed_dict):
        return self.run([self.G_optim, self.G_loss], feed_dict)[1]
