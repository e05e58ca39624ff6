"""Drop-in surface of the reference's ``models/stackgan/stageI/model.py`` (class ConditionalGan, :5-171) on the
B200-native engine.  Same constructor, the same attribute names for the inputs
(``inputs, wrong_inputs, embed_inputs, z, z_sample, embed_sample``), the fetchable tensors (``G, D_synthetic_logits,
D_real_match_logits, D_real_mismatch_logits, sampler``), ``d_vars / g_vars`` and the ``generator`` /
``discriminator`` signatures.  The losses and optimizers are declared by the trainer in the reference
(models/stackgan/stageI/trainer.py:19-55); ``ConditionalGanTrainer.define_losses`` of this package declares the same
names and executes them through ``model.run(fetches, feed_dict)``, the stand-in for ``sess.run``.

All arithmetic runs in the CUDA library through ``t2i_b200.kernels`` (engine_stage1.StageIEngine); there is no CPU
fallback.
"""
import numpy as np
import torch

from ....engine_stage1 import StageIEngine
from ...wgancls.model import Fetch, Placeholder, _truncated_normal


class ConditionalGan(object):
    def __init__(self, cfg, build_model=True, precision="bf16", device=None, kernels=None, distributed=None,
                 use_graphs=True, sync_bn=False):
        """
        Args:
          cfg: Config specifying all the parameters of the model (reference: model.py:6-10).
          precision / device / kernels / distributed / sync_bn: as for WGanCls (models/wgancls/model.py of this
            package); sync_bn covers the BatchNorms of BOTH networks (d_net has one after almost every conv).
        """
        self.name = 'ConditionalGAN/StageI'
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
            # the reference hard-codes the 4x4 embedding tile (model.py:106): only 64x64x3 is valid
            raise ValueError("StackGAN stage-I is defined for 64x64x3 images only (got %s)" % (self.image_dims,))

        self.precision = precision
        self._np = {"bf16": 1, "bf16x3": 2}[precision]
        if kernels is None:
            from .... import kernels as _k      # loads libt2i_b200.so lazily, raises if missing
            kernels = _k
        self._K = kernels
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("ConditionalGan needs a CUDA device (sm_100a); there is no CPU fallback")
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
        self._train_engine()
        if build_model:
            self.build_model()

    def _engine(self, batch):
        if batch not in self._engines:
            base = next(iter(self._engines.values()), None)
            t = self.cfg.TRAIN
            self._engines[batch] = StageIEngine(
                self._K, self.device, batch, self._np, self.z_dim, self.embed_dim, self.compressed_embed_dim,
                self.gf_dim, self.df_dim, t.D_BETA_DECAY, t.G_BETA_DECAY, t.COEFF.ALPHA_MISMATCH_LOSS, t.COEFF.KL,
                self._world, self._allreduce, share_from=base, use_graphs=self._use_graphs, sync_bn=self._sync_bn)
        return self._engines[batch]

    def _train_engine(self):
        return self._engine(self.batch_size)

    def build_model(self):
        """model.py:36-56: the placeholders and the fetchable tensors of the reference graph, by name."""
        b, dims = self.batch_size, self.image_dims
        self.inputs = Placeholder("real_images", [b] + dims)
        self.wrong_inputs = Placeholder("wrong_images", [b] + dims)
        self.embed_inputs = Placeholder("phi_inputs", [b, self.embed_dim])
        self.z = Placeholder("z", [b, self.z_dim])
        self.z_sample = Placeholder("z_sample", [self.sample_num, self.z_dim])
        self.embed_sample = Placeholder("phi_sample", [self.sample_num, self.embed_dim])
        # explicit noise inputs (the reference draws them inside the graph, model.py:71); optional
        self.cond_noise = Placeholder("cond_noise", [b, self.compressed_embed_dim])
        self.cond_noise_sample = Placeholder("cond_noise_sample", [self.sample_num, self.compressed_embed_dim])
        for n in ("G", "embed_mean", "embed_log_sigma", "D_synthetic", "D_synthetic_logits", "D_real_match",
                  "D_real_match_logits", "D_real_mismatch", "D_real_mismatch_logits", "sampler"):
            setattr(self, n, Fetch(n, "tensor"))
        self.d_vars = [n for n in self.variable_names() if n.startswith("d_net/") and "moving_" not in n]
        self.g_vars = [n for n in self.variable_names() if n.startswith("g_net/") and "moving_" not in n]
        self._built = True

    # ------------------------------------------------------------------ variables (checkpoint boundary)
    def variable_names(self):
        return list(self._train_engine().get_params_tf().keys())

    def get_variables(self):
        return self._train_engine().get_params_tf()

    def set_variables(self, variables):
        self._train_engine().set_params_tf(dict(variables))

    def get_optimizer_state(self):
        return self._train_engine().get_adam_tf()

    def set_optimizer_state(self, state):
        if state:
            self._train_engine().set_adam_tf(state)

    def initialize(self, seed=0):
        """tf.global_variables_initializer(): w_init = N(0, 0.02) for every conv / deconv / generator dense kernel
        (model.py:29), gamma ~ N(1, 0.02) (:30-32), zero biases / beta, moving statistics 0 / 1.  (The discriminator's
        embedding dense keeps TF's glorot default in the reference; N(0, 0.02) is used for it here as well.)"""
        gen = torch.Generator().manual_seed(seed)
        eng = self._train_engine()
        p = eng.get_params_tf()
        for name, w in p.items():
            leaf = name.rsplit("/", 1)[1]
            if leaf in ("weights", "kernel"):
                p[name] = torch.randn(w.shape, generator=gen) * 0.02
            elif leaf == "gamma":
                p[name] = 1.0 + 0.02 * torch.randn(w.shape, generator=gen)
            elif leaf == "moving_variance":
                p[name] = torch.ones_like(w)
            else:
                p[name] = torch.zeros_like(w)
        eng.set_params_tf(p)
        for k in ("d", "g"):
            eng.adam_m[k].zero_()
            eng.adam_v[k].zero_()
        eng.d_t = eng.g_t = 0

    # ------------------------------------------------------------------ eager sub-graphs
    def _dev(self, a, shape=None):
        t = torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a)
        t = t.to(self.device, torch.float32, non_blocking=True)
        return t.reshape(shape) if shape is not None else t

    def seed_noise(self, seed):
        self._noise_gen = torch.Generator(device=self.device).manual_seed(seed)

    def generator(self, z, embed, is_training=True, reuse=False, cond_noise=True, noise=None):
        """model.py:114-171.  Returns (image NHWC in (-1,1), mean, log_sigma) as CUDA fp32 tensors."""
        z = self._dev(z)
        b = z.shape[0]
        eng = self._engine(b)
        embed = self._dev(embed, (b, self.embed_dim))
        tn = self._dev(noise) if noise is not None else _truncated_normal((b, self.compressed_embed_dim),
                                                                         self.device, self._noise_gen)
        out = torch.empty(b, 64, 64, 3, device=self.device, dtype=torch.float32)
        zp = torch.zeros(b, eng.Z, device=self.device, dtype=torch.float32)
        zp[:, :self.z_dim] = z
        eng.g["kl_scratch"].zero_()
        eng.g_forward(zp, embed, tn, out, eng.g["kl_scratch"], train=is_training, cond_noise=cond_noise)
        ms = eng.g["ms"].clone()        # fp32 [b, 2*ce] = [mean | log_sigma]
        ce = self.compressed_embed_dim
        return out, ms[:, :ce], ms[:, ce:]

    def discriminator(self, inputs, embed, is_training=True, reuse=False):
        """model.py:76-112.  inputs NHWC [B,64,64,3], embed [B,1024] -> (sigmoid(logits), logits), each [B,1,1,1].
        BatchNorm uses the batch statistics of THIS call, as every call site of the reference does (:45-49)."""
        if not is_training:
            raise NotImplementedError("the reference never builds d_net with is_training=False")
        x = self._dev(inputs)
        b = x.shape[0]
        eng = self._engine(b)
        logits = eng.discriminator_logits(x, self._dev(embed, (b, self.embed_dim))).reshape(b, 1, 1, 1)
        return torch.sigmoid(logits), logits

    # ------------------------------------------------------------------ sess.run stand-in
    def run(self, fetches, feed_dict=None, lr=None):
        """``sess.run(fetches, feed_dict)`` for the trainer's fetch lists: [D_optim, D_loss, ...] (trainer.py:139),
        [G_optim, G_loss, ...] (:144), sampler (:156).  Train ops return None, scalars floats, tensors numpy."""
        if not self._built:
            raise RuntimeError("the model was constructed with build_model=False")
        feed_dict = feed_dict or {}
        single = not isinstance(fetches, (list, tuple))
        flist = [fetches] if single else list(fetches)
        names = [f.name for f in flist]
        eng = self._train_engine()
        if lr is None:
            lr = next((v for k, v in feed_dict.items() if getattr(k, "name", None) == "lr"), self.cfg.TRAIN.D_LR)
        b = self.batch_size
        t = lambda a: None if a is None else torch.as_tensor(np.asarray(a, dtype=np.float32) if not torch.is_tensor(a) else a)
        get = lambda k: feed_dict[k] if k in feed_dict else None
        ran = False
        for op, step in (("D_optim", eng.d_step), ("G_optim", eng.g_step)):
            if op in names:
                noise = get(self.cond_noise)
                if noise is None:      # a fresh truncated-normal draw per run (model.py:71)
                    noise = _truncated_normal((b, self.compressed_embed_dim), self.device, self._noise_gen)
                eng.load_feed(x=t(get(self.inputs)), x_mismatch=t(get(self.wrong_inputs)), cond=t(get(self.embed_inputs)),
                              z=t(get(self.z)), tn_eps=t(noise))
                step(float(lr))
                ran = True
        out, sc = [], None
        for f in flist:
            if f.kind == "op":
                out.append(None)
            elif f.kind == "scalar":
                if not ran and sc is None:
                    raise RuntimeError("scalar '%s' is produced by the D/G run; fetch it with the train op" % f.name)
                sc = sc or eng.scalars_dict()
                out.append(sc[f.name])
            elif f.name == "sampler":
                img, _, _ = self.generator(feed_dict[self.z_sample], feed_dict[self.embed_sample], is_training=False,
                                           noise=get(self.cond_noise_sample))
                out.append(img.cpu().numpy())
            elif f.name == "G":
                out.append(eng.d["img"][:b].cpu().numpy())
            elif f.name in ("D_synthetic_logits", "D_real_match_logits", "D_real_mismatch_logits"):
                k = ["D_synthetic_logits", "D_real_match_logits", "D_real_mismatch_logits"].index(f.name)
                out.append(eng.d["logit"][k * b:(k + 1) * b].cpu().numpy().reshape(b, 1, 1, 1))
            else:
                raise KeyError("fetch '%s' is not materialised by this implementation" % f.name)
        return out[0] if single else out
