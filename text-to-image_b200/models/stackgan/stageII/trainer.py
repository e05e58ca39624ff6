"""Mirror of the reference's ``models/stackgan/stageII/trainer.py`` (ConditionalGanTrainer, :11-175) on the B200
model: same constructor (``sess`` accepted and ignored; ``cfg_stage_i`` names the stage-I checkpoint directory),
``define_losses`` declares the reference's loss / optimizer names and the two savers (:20-59), ``train`` keeps its
step order (D run then G run on the same feed, :131-137), learning-rate schedule (lr * 0.5 ** (epoch // 100),
:113,121), the two-checkpoint restore (stage-II networks, then the stage-I generator, :98-109), sampling every
2000 updates (:160-173), a stage-II checkpoint when counter % 500 == 2 (:175-176).  Histogram / image summaries and PNG writing depend on TF and removed scipy APIs and are out
of scope; loss scalars go to ``self.log``."""
import time

import numpy as np

from ...wgancls.model import Fetch, Placeholder
from ....utils.saver import load, save


class _ScopeSaver(object):
    """tf.train.Saver(var_list=<variables of some scopes>) (trainer.py:48-52): the save / load helpers of
    utils/saver.py see only the variables under ``prefixes``; optimizer slots travel with the trainable scopes."""

    def __init__(self, model, prefixes, with_optimizer):
        self.model, self.prefixes, self.with_optimizer = model, tuple(prefixes), with_optimizer

    def get_variables(self):
        return {k: v for k, v in self.model.get_variables().items() if k.startswith(self.prefixes)}

    def set_variables(self, variables):
        self.model.set_variables({k: v for k, v in variables.items() if k.startswith(self.prefixes)})

    def get_optimizer_state(self):
        return self.model.get_optimizer_state() if self.with_optimizer else {}

    def set_optimizer_state(self, state):
        if self.with_optimizer:
            self.model.set_optimizer_state(state)


class ConditionalGanTrainer(object):
    def __init__(self, sess, model, dataset, cfg, cfg_stage_i, log=None, on_samples=None):
        self.sess = sess            # ignored: there is no session, the model runs itself
        self.model = model
        self.dataset = dataset
        self.cfg = cfg
        self.cfg_stage_i = cfg_stage_i
        self.lr = self.cfg.TRAIN.D_LR
        self.log = log if log is not None else []
        self.on_samples = on_samples

    def define_losses(self):
        """trainer.py:20-59: names of the losses, the two savers and the two train ops."""
        self.learning_rate = Placeholder("lr")
        for n in ("D_synthetic_loss", "D_real_match_loss", "D_real_mismatch_loss", "G_kl_loss", "G_gan_loss", "D_loss",
                  "G_loss"):
            setattr(self, n, Fetch(n, "scalar"))
        self.stagei_g_saver = _ScopeSaver(self.model, ("g_net/",), False)
        self.stageii_saver = _ScopeSaver(self.model, ("stageII_g_net/", "stageII_d_net/"), True)
        self.D_optim = Fetch("D_optim", "op")
        self.G_optim = Fetch("G_optim", "op")

    def train(self, max_updates=None):
        self.define_losses()
        m, cfg = self.model, self.cfg
        sample_z = np.random.normal(0, 1, (m.sample_num, m.z_dim))
        _, sample_embed, _, captions = self.dataset.test.next_batch_test(m.sample_num, 0, 1)
        sample_embed = np.squeeze(sample_embed, axis=0)
        print(sample_embed.shape)

        counter = 1
        start_time = time.time()
        m.initialize()       # initialize_uninitialized (trainer.py:111): whatever the two loads below do not restore

        could_load, checkpoint_counter = load(self.stageii_saver, cfg.CHECKPOINT_DIR, prefix="stageII")
        if could_load:
            counter = checkpoint_counter
            print(" [*] Load SUCCESS: Stage II networks are loaded.")
        else:
            print(" [!] Load failed for stage II networks...")

        could_load, checkpoint_counter = load(self.stagei_g_saver, self.cfg_stage_i.CHECKPOINT_DIR, prefix="stageI")
        if could_load:
            print(" [*] Load SUCCESS: Stage I generator is loaded")
        else:
            print(" [!] WARNING!!! Failed to load the parameters for stage I generator...")

        updates_per_epoch = self.dataset.train.num_examples // m.batch_size
        epoch_start = counter // updates_per_epoch
        done = 0
        for epoch in range(epoch_start, cfg.TRAIN.EPOCH):
            cen_epoch = epoch // 100
            for idx in range(0, updates_per_epoch):
                images, wrong_images, embed, _, _ = self.dataset.train.next_batch(m.batch_size, 4, embeddings=True,
                                                                                  wrong_img=True)
                batch_z = np.random.normal(0, 1, (m.batch_size, m.z_dim))
                feed_dict = {
                    self.learning_rate: self.lr * (0.5 ** cen_epoch),
                    m.inputs: images,
                    m.wrong_inputs: wrong_images,
                    m.embed_inputs: embed,
                    m.z: batch_z,
                }
                # Update D network, then G network, on the same feed (trainer.py:131-137)
                _, err_d = m.run([self.D_optim, self.D_loss], feed_dict=feed_dict)
                _, err_g = m.run([self.G_optim, self.G_loss], feed_dict=feed_dict)
                counter += 1
                self.log.append({"epoch": epoch, "idx": idx, "counter": counter, "d_loss": err_d, "g_loss": err_g,
                                 "lr": feed_dict[self.learning_rate], "time": time.time() - start_time})
                if np.mod(counter, 2000) == 0:
                    samples = m.run(m.sampler, feed_dict={m.z_sample: sample_z, m.embed_sample: sample_embed})
                    if self.on_samples is not None:
                        self.on_samples(epoch, idx, samples, captions)
                if np.mod(counter, 500) == 2:
                    save(self.stageii_saver, cfg.CHECKPOINT_DIR, counter, cfg.TRAIN.CHECKPOINTS_TO_KEEP, prefix="stageII")
                done += 1
                if max_updates is not None and done >= max_updates:
                    return
