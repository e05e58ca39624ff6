"""Drop-in surface of the reference's ``models/stackgan/stageII/model.py`` (class ConditionalGan, :7-201) on the
B200-native engine.  Same constructor ``ConditionalGan(stagei, cfg, build_model=True)`` --
``stagei`` is the stage-I model object whose generator (variables ``g_net/*``) runs, frozen, inside the stage-II
graph (:50-51) --, the same attribute names for the inputs (``inputs, wrong_inputs, embed_inputs, z, z_sample,
embed_sample``), the fetchable tensors (``G, D_synthetic_logits, D_real_match_logits, D_real_mismatch_logits,
sampler``), ``d_vars / g_vars`` and the ``generator(image, embed, ...)`` / ``discriminator(inputs, embed, ...)``
signatures.  Losses and optimizers are declared by the trainer (models/stackgan/stageII/trainer.py:20-59) and executed
through ``model.run(fetches, feed_dict)``, the stand-in for ``sess.run``.

All arithmetic runs in the CUDA library through ``t2i_b200.kernels`` (engine_stage2.StageIIEngine); there is no CPU
fallback.
"""
import numpy as np
import torch

from ....engine_stage2 import D2, G2, StageIIEngine
from ...wgancls.model import Fetch, Placeholder, _truncated_normal


class ConditionalGan(object):
    def __init__(self, stagei, cfg, build_model=True, precision=None, device=None, kernels=None, distributed=None,
                 use_graphs=True, sync_bn=False):
        """
        Args:
          stagei: the stage-I ``ConditionalGan`` of this package (models/stackgan/stageI/model.py); its generator
            parameters are shared with the stage-II engine, exactly one copy exists.
          cfg: Config specifying all the parameters of the model (reference: model.py:8-12).
          precision / device / kernels: default to those of ``stagei``.
          distributed / sync_bn: as for WGanCls; sync_bn makes all three networks (frozen stage-I generator, stageII_g_net,
            stageII_d_net) normalise with the statistics of the GLOBAL batch, as the single-device reference does.
        """
        self.name = 'ConditionalGAN/StageII'
        self.stagei = stagei
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
        if self.output_size != 256 or self.image_dims != [256, 256, 3]:
            # the generator always upsamples 16 -> 256 (model.py:160-178): any other size breaks the graph
            raise ValueError("StackGAN stage-II is defined for 256x256x3 images only (got %s)" % (self.image_dims,))
        for a in ("z_dim", "embed_dim", "compressed_embed_dim"):
            if getattr(stagei, a) != getattr(self, a):
                raise ValueError("stage-I and stage-II configs disagree on %s" % a)

        self.precision = precision or stagei.precision
        if self.precision != stagei.precision:
            raise ValueError("stage-I and stage-II must use the same precision (shared generator parameters)")
        self._np = {"bf16": 1, "bf16x3": 2}[self.precision]
        self._K = kernels if kernels is not None else stagei._K
        self.device = torch.device(device) if device is not None else stagei.device
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
            self._engines[batch] = StageIIEngine(
                self._K, self.device, batch, self._np, self.z_dim, self.embed_dim, self.compressed_embed_dim,
                self.gf_dim, self.df_dim, t.D_BETA_DECAY, t.G_BETA_DECAY, t.COEFF.ALPHA_MISMATCH_LOSS, t.COEFF.KL,
                self._world, self._allreduce, s1_engine=self.stagei._train_engine(), share_from=base,
                use_graphs=self._use_graphs, sync_bn=self._sync_bn)
        return self._engines[batch]

    def _train_engine(self):
        return self._engine(self.batch_size)

    def build_model(self):
        """model.py:40-60: the placeholders and the fetchable tensors of the reference graph, by name."""
        b, dims = self.batch_size, self.image_dims
        self.inputs = Placeholder("real_images", [b] + dims)
        self.wrong_inputs = Placeholder("wrong_images", [b] + dims)
        self.embed_inputs = Placeholder("phi_inputs", [b, self.embed_dim])
        self.z = Placeholder("z", [b, self.z_dim])
        self.z_sample = Placeholder("z_sample", [self.sample_num, self.z_dim])
        self.embed_sample = Placeholder("phi_sample", [self.sample_num, self.embed_dim])
        # explicit noise inputs (the reference draws them inside the graph: stage-I model.py:71, stage-II :73); optional
        ce = self.compressed_embed_dim
        self.cond_noise = Placeholder("cond_noise", [b, ce])
        self.cond_noise_stagei = Placeholder("cond_noise_stagei", [b, ce])
        self.cond_noise_sample = Placeholder("cond_noise_sample", [self.sample_num, ce])
        self.cond_noise_stagei_sample = Placeholder("cond_noise_stagei_sample", [self.sample_num, ce])
        for n in ("G", "embed_mean", "embed_log_sigma", "D_synthetic", "D_synthetic_logits", "D_real_match",
                  "D_real_match_logits", "D_real_mismatch", "D_real_mismatch_logits", "sampler"):
            setattr(self, n, Fetch(n, "tensor"))
        self.d_vars = [n for n in self.variable_names() if n.startswith(D2) and "moving_" not in n]
        self.g_vars = [n for n in self.variable_names() if n.startswith(G2) and "moving_" not in n]
        self._built = True

    # ------------------------------------------------------------------ variables (checkpoint boundary)
    def variable_names(self):
        return list(self._train_engine().get_params_tf().keys())

    def get_variables(self):
        """stageII_g_net/*, stageII_d_net/* and the shared stage-I g_net/* (trainer.py:48-51 saves them separately)."""
        return self._train_engine().get_params_tf()

    def set_variables(self, variables):
        """A partial dict is merged into the current values (the two savers of trainer.py:50-52 each restore a
        subset: stage-II networks from one directory, the stage-I generator from another)."""
        eng = self._train_engine()
        cur = eng.get_params_tf()
        unknown = [k for k in variables if k not in cur and not k.startswith("d_net/")]
        if unknown:
            raise KeyError("unknown variables: %s" % unknown[:4])
        cur.update({k: torch.as_tensor(np.asarray(v)) for k, v in variables.items() if k in cur})
        eng.set_params_tf(cur)

    def get_optimizer_state(self):
        return self._train_engine().get_adam_tf()

    def set_optimizer_state(self, state):
        if state:
            self._train_engine().set_adam_tf(state)

    def initialize(self, seed=0):
        """tf.global_variables_initializer() for the stage-II scopes: N(0, 0.02) kernels (model.py:31), gamma ~
        N(1, 0.02) (:32-34), zero biases / beta, moving statistics 0 / 1.  The stage-I generator is left alone."""
        gen = torch.Generator().manual_seed(seed)
        eng = self._train_engine()
        p = eng.get_params_tf()
        for name, w in p.items():
            if name.startswith("g_net/"):
                continue
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

    def _noise(self, given, b):
        if given is not None:
            return self._dev(given)
        return _truncated_normal((b, self.compressed_embed_dim), self.device, self._noise_gen)

    def generator(self, image, embed, is_training=True, reuse=False, cond_noise=True, noise=None):
        """model.py:180-201.  image: a stage-I output, NHWC [B,64,64,3] -> (image NHWC [B,256,256,3] in (-1,1), mean,
        log_sigma) as CUDA fp32 tensors."""
        image = self._dev(image)
        b = image.shape[0]
        eng = self._engine(b)
        eng.feed["cond"].copy_(self._dev(embed, (b, self.embed_dim)))
        eng.g["img64"].copy_(image)
        eng.g["tn"].copy_(self._noise(noise, b))
        out = torch.empty(b, 256, 256, 3, device=self.device, dtype=torch.float32)
        eng.g["kl_scratch"].zero_()
        eng.g2_forward(out, eng.g["kl_scratch"], train=is_training, cond_noise=cond_noise, update_moving=False,
                       run_stage1=False)
        ce = self.compressed_embed_dim
        ms = eng.g["ms"].clone()        # fp32 [b, 2*ce] = [mean | log_sigma]
        return out, ms[:, :ce], ms[:, ce:]

    def discriminator(self, inputs, embed, is_training=True, reuse=False):
        """model.py:78-132.  inputs NHWC [B,256,256,3], embed [B,1024] -> (sigmoid(logits), logits), each [B,1,1,1].
        BatchNorm uses the batch statistics of THIS call, as every call site of the reference does (:53-56)."""
        if not is_training:
            raise NotImplementedError("the reference never builds stageII_d_net with is_training=False")
        x = self._dev(inputs)
        b = x.shape[0]
        eng = self._engine(b)
        logits = eng.discriminator_logits(x, self._dev(embed, (b, self.embed_dim))).reshape(b, 1, 1, 1)
        return torch.sigmoid(logits), logits

    def sample(self, z, embed, noise_stagei=None, noise=None):
        """The ``sampler`` tensor (model.py:51,57): stage-I and stage-II generators with is_training=False."""
        z = self._dev(z)
        b = z.shape[0]
        eng = self._engine(b)
        eng.load_feed(cond=self._dev(embed, (b, self.embed_dim)), z=z, tn_eps=self._noise(noise, b),
                      tn_s1=self._noise(noise_stagei, b))
        out = torch.empty(b, 256, 256, 3, device=self.device, dtype=torch.float32)
        eng.g["kl_scratch"].zero_()
        eng.g2_forward(out, eng.g["kl_scratch"], train=False, update_moving=False)
        return out

    # ------------------------------------------------------------------ sess.run stand-in
    def run(self, fetches, feed_dict=None, lr=None):
        """``sess.run(fetches, feed_dict)`` for the trainer's fetch lists: [D_optim, D_loss, ...] (trainer.py:131),
        [G_optim, G_loss, ...] (:136), sampler (:147).  Train ops return None, scalars floats, tensors numpy."""
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
            if op in names:     # a fresh truncated-normal draw per run and per generator (stage-I :71, stage-II :73)
                img = op == "D_optim"       # the G run reads no real image (its graph ends at D(G(z)), trainer.py:136)
                eng.load_feed(x=t(get(self.inputs)) if img else None, x_mismatch=t(get(self.wrong_inputs)) if img else None,
                              cond=t(get(self.embed_inputs)),
                              z=t(get(self.z)), tn_eps=t(self._noise(get(self.cond_noise), b)),
                              tn_s1=t(self._noise(get(self.cond_noise_stagei), b)))
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
                img = self.sample(feed_dict[self.z_sample], feed_dict[self.embed_sample],
                                  get(self.cond_noise_stagei_sample), get(self.cond_noise_sample))
                out.append(img.cpu().numpy())
            elif f.name == "G":
                out.append(eng.d["img"][:b].cpu().numpy())
            elif f.name in ("D_synthetic_logits", "D_real_match_logits", "D_real_mismatch_logits"):
                k = ["D_synthetic_logits", "D_real_match_logits", "D_real_mismatch_logits"].index(f.name)
                out.append(eng.d["logit"][k * b:(k + 1) * b].cpu().numpy().reshape(b, 1, 1, 1))
            else:
                raise KeyError("fetch '%s' is not materialised by this implementation" % f.name)
        return out[0] if single else out
