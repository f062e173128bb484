"""`FDReturn` record (learner/fd_return.py:5-23): what a worker hands the learner."""


class FDReturn(object):
    def __init__(self):
        self.epoch = -1
        self.encoded_noise = "-1"
        self.perturbation = None
        self.reward = 0
        self.novelty = 0
        self.entropy = 0
        self.timesteps = 0
        self.is_eval = False
        self.eval_states = []
        self.obs_stats_update = []

    def serialize(self):
        return self.reward, self.novelty, self.entropy, self.timesteps, self.encoded_noise, self.perturbation, \
               self.epoch, self.is_eval, self.eval_states, self.obs_stats_update

    def deserialize(self, other):
        self.reward, self.novelty, self.entropy, self.timesteps, self.encoded_noise, self.perturbation, self.epoch, \
            self.is_eval, self.eval_states, self.obs_stats_update = other
