"""SAVi / AVLEN policies (ss_baselines/savi/ppo/policy.py:39-674) behind the reference's API, on CUDA kernels.

``Policy.act / get_value / evaluate_actions`` keep the reference signatures, argument meaning and return order
(SURVEY.md §8b).  Parameter names and shapes equal the reference's ``state_dict`` (Appendix A) so checkpoints load.
"""
from __future__ import annotations

import abc
import itertools
import logging

import torch
import torch.nn as nn

from ... import _lib
from ... import nn as K
from ...common.utils import CategoricalNet, cuda_linear
from ..models.audio_cnn import AudioCNN
from ..models.clip_text import CLIPTextTower
from ..models.dialog_state_encoder import DialogStateEncoder
from ..models.smt_cnn import SMTCNN
from ..models.smt_state_encoder import IndexedMemory, SMTStateEncoder

SPECTROGRAM, POSE, CATEGORY = "spectrogram", "pose", "category"
CATEGORY_BELIEF, LOCATION_BELIEF = "category_belief", "location_belief"


class CriticHead(nn.Module):
    def __init__(self, input_size):
        super().__init__()
        self.fc = nn.Linear(input_size, 1)
        nn.init.orthogonal_(self.fc.weight)
        nn.init.constant_(self.fc.bias, 0)

    def forward(self, x):
        return cuda_linear(x, self.fc.weight, self.fc.bias)


class CriticHead2(nn.Module):
    def __init__(self, input_size):
        super().__init__()
        self.fc = nn.Linear(input_size, 2)
        nn.init.orthogonal_(self.fc.weight)
        nn.init.constant_(self.fc.bias, 0)

    def forward(self, x):
        return cuda_linear(x, self.fc.weight, self.fc.bias)


class Policy(nn.Module):
    """policy.py:39-276.  Every policy owns all seven heads, used or not (Appendix A)."""

    def __init__(self, net, dim_actions, dim_actions_option=2):
        super().__init__()
        self.net = net
        self.dim_actions = dim_actions
        self.dim_actions_option = dim_actions_option
        self.action_distribution_option = CategoricalNet(self.net.output_size, self.dim_actions_option)
        self.action_distribution_goal = CategoricalNet(self.net.output_size, self.dim_actions)
        self.action_distribution_vln = CategoricalNet(self.net.output_size, self.dim_actions)
        self.critic_goal = CriticHead(self.net.output_size)
        self.critic_option = CriticHead(self.net.output_size)
        self.uncertainty_option = CriticHead2(self.net.output_size)
        self.critic_vln = CriticHead(self.net.output_size)

    def forward(self, *x):
        raise NotImplementedError

    def act(self, observations, rnn_hidden_states, prev_actions, masks, ext_memory, ext_memory_masks,
            deterministic=False, uniforms=None):
        features, rnn_hidden_states, ext_memory_feats = self.net(
            observations, rnn_hidden_states, prev_actions, masks, ext_memory, ext_memory_masks)
        distribution, _ = self.action_distribution_goal(features)
        value = self.critic_goal(features)
        action = distribution.mode() if deterministic else distribution.sample(uniforms=uniforms)
        action_log_probs = distribution.log_probs(action)
        return value, action, action_log_probs, rnn_hidden_states, ext_memory_feats, distribution.probs

    def act_option(self, observations, rnn_hidden_states, prev_actions, masks, ext_memory, ext_memory_masks,
                   query_state, last_query_info, deterministic=False, uniforms=None):
        features, rnn_hidden_states, ext_memory_feats = self.net(
            observations, rnn_hidden_states, prev_actions, masks, ext_memory, ext_memory_masks, query_state,
            last_query_info)
        distribution, _ = self.action_distribution_option(features)
        value = self.critic_option(features)
        unct = self.uncertainty_option(features)
        action = distribution.mode() if deterministic else distribution.sample(uniforms=uniforms)
        action_log_probs = distribution.log_probs(action)
        return value, unct, action, action_log_probs, rnn_hidden_states, ext_memory_feats, distribution.probs

    def act_dialog(self, observations, rnn_hidden_states, prev_actions, masks, ext_memory, ext_memory_dialog,
                   ext_memory_masks, all_dialog, agent_step, deterministic=False, without_dialog=False, uniforms=None,
                   scene=None):
        """policy.py:130-162 (pi_l).  ``scene``: optional result of ``net.encode_scene`` computed ahead (side stream)."""
        if without_dialog:
            all_dialog = None
        kw = {} if scene is None else {"scene": scene}
        features, rnn_hidden_states, ext_memory_feats, ext_memory_dialog_feats = self.net(
            observations, rnn_hidden_states, prev_actions, masks, ext_memory, ext_memory_dialog, ext_memory_masks,
            all_dialog, agent_step, **kw)
        distribution, _ = self.action_distribution_vln(features)
        value = self.critic_vln(features)
        action = distribution.mode() if deterministic else distribution.sample(uniforms=uniforms)
        action_log_probs = distribution.log_probs(action)
        return (value, action, action_log_probs, rnn_hidden_states, ext_memory_feats, ext_memory_dialog_feats,
                distribution.probs)

    def get_value(self, observations, rnn_hidden_states, prev_actions, masks, ext_memory, ext_memory_masks):
        features, _, _ = self.net(observations, rnn_hidden_states, prev_actions, masks, ext_memory, ext_memory_masks)
        return self.critic_goal(features)

    def get_value_option(self, observations, rnn_hidden_states, prev_actions, masks, ext_memory, ext_memory_masks,
                         query_state, last_query_info):
        features, _, _ = self.net(observations, rnn_hidden_states, prev_actions, masks, ext_memory, ext_memory_masks,
                                  query_state, last_query_info)
        return self.critic_option(features)

    def evaluate_actions(self, observations, rnn_hidden_states, prev_actions, masks, action, ext_memory,
                         ext_memory_masks):
        features, rnn_hidden_states, ext_memory_feats = self.net(
            observations, rnn_hidden_states, prev_actions, masks, ext_memory, ext_memory_masks)
        distribution, _ = self.action_distribution_goal(features)
        value = self.critic_goal(features)
        action_log_probs = distribution.log_probs(action)
        distribution_entropy = distribution.entropy().mean()
        return value, action_log_probs, distribution_entropy, rnn_hidden_states, ext_memory_feats

    def evaluate_actions_option(self, observations, rnn_hidden_states, prev_actions, masks, action, ext_memory,
                                ext_memory_masks, query_state, last_query_info):
        features, rnn_hidden_states, ext_memory_feats = self.net(
            observations, rnn_hidden_states, prev_actions, masks, ext_memory, ext_memory_masks, query_state,
            last_query_info)
        distribution, _ = self.action_distribution_option(features)
        value = self.critic_option(features)
        unct = self.uncertainty_option(features)
        action_log_probs = distribution.log_probs(action)
        distribution_entropy = distribution.entropy().mean()
        return (value, unct, action_log_probs, distribution_entropy, rnn_hidden_states, ext_memory_feats,
                distribution.probs)

    def evaluate_actions_dialog(self, observations, rnn_hidden_states, prev_actions, masks, action, ext_memory,
                                ext_memory_dialog, ext_memory_masks, all_dialog, agent_step, without_dialog=False):
        """policy.py:238-276: value is None; the raw logits are returned for the imitation loss (ppo.py:123-132)."""
        if without_dialog:
            all_dialog = None
        features, rnn_hidden_states, ext_memory_feats, ext_memory_dialog_feats = self.net(
            observations, rnn_hidden_states, prev_actions, masks, ext_memory, ext_memory_dialog, ext_memory_masks,
            all_dialog, agent_step)
        distribution, logit = self.action_distribution_vln(features)
        action_log_probs = distribution.log_probs(action)
        distribution_entropy = distribution.entropy().mean()
        return (None, action_log_probs, distribution_entropy, rnn_hidden_states, ext_memory_feats,
                ext_memory_dialog_feats, logit)

    # ---- fused path used by PPO.update: features -> raw head outputs, no distribution objects -------------
    def evaluate_heads(self, which, observations, rnn_hidden_states, prev_actions, masks, ext_memory,
                       ext_memory_masks, *extra):
        features, _, _ = self.net(observations, rnn_hidden_states, prev_actions, masks, ext_memory, ext_memory_masks,
                                  *extra)
        if which == "goal":
            head, critic = self.action_distribution_goal, self.critic_goal
        elif which == "option":
            head, critic = self.action_distribution_option, self.critic_option
        else:
            head, critic = self.action_distribution_vln, self.critic_vln
        logits = cuda_linear(features, head.linear.weight, head.linear.bias)
        value = critic(features)
        unct = self.uncertainty_option(features) if which == "option" else None
        return logits, value, unct


class Net(nn.Module, metaclass=abc.ABCMeta):
    @abc.abstractmethod
    def forward(self, observations, rnn_hidden_states, prev_actions, masks):
        pass

    @property
    @abc.abstractmethod
    def output_size(self):
        pass

    @property
    @abc.abstractmethod
    def num_recurrent_layers(self):
        pass

    @property
    @abc.abstractmethod
    def is_blind(self):
        pass


class AudioNavSMTNet(Net):
    """policy.py:501-674: SMTCNN(rgb, depth) | action embedding | AudioCNN(spectrogram) | [category] | pose
    -> scene-memory transformer with the belief vector as decoder query."""

    def __init__(self, observation_space, action_space, hidden_size=128, use_pretrained=False, pretrained_path="",
                 use_belief_as_goal=True, use_label_belief=True, use_location_belief=True, use_belief_encoding=False,
                 normalize_category_distribution=False, use_category_input=False, **kwargs):
        super().__init__()
        self._use_action_encoding = True
        self._use_residual_connection = False
        self._use_belief_as_goal = use_belief_as_goal
        self._use_label_belief = use_label_belief
        self._use_location_belief = use_location_belief
        self._hidden_size = hidden_size
        self._action_size = action_space.n
        self._use_belief_encoder = use_belief_encoding
        self._normalize_category_distribution = normalize_category_distribution
        self._use_category_input = use_category_input
        if not use_belief_as_goal or use_belief_encoding:
            raise _lib.AvlenError("only use_belief_as_goal=True, use_belief_encoding=False is built "
                                  "(the setting of every reference yaml)")
        assert SPECTROGRAM in observation_space.spaces
        self.goal_encoder = AudioCNN(observation_space, 128, SPECTROGRAM)
        audio_feature_dims = 128
        self.visual_encoder = SMTCNN(observation_space)
        self.action_encoder = nn.Linear(self._action_size, 16)
        nfeats = self.visual_encoder.feature_dims + 16 + audio_feature_dims
        self._cat_col = nfeats
        if self._use_category_input:
            nfeats += 21
        assert POSE in observation_space.spaces
        pose_dims = observation_space.spaces[POSE].shape[0]
        pose_indices = (nfeats, nfeats + pose_dims)
        nfeats += pose_dims
        self._base_feature_size = nfeats
        nfeats += self._extra_feature_dims()
        self._feature_size = nfeats
        self._build_extra()
        self.smt_state_encoder = SMTStateEncoder(nfeats, dim_feedforward=hidden_size, pose_indices=pose_indices,
                                                 **kwargs)
        self._post_init(kwargs)
        self.state_size = self.smt_state_encoder.hidden_state_size
        if use_pretrained:
            assert pretrained_path != ""
            self.pretrained_initialization(pretrained_path)
        self.train()

    def _extra_feature_dims(self):
        return 0

    def _build_extra(self):
        pass

    def _post_init(self, kwargs):
        pass

    @property
    def memory_dim(self):
        return self._feature_size

    @property
    def output_size(self):
        return self.smt_state_encoder.hidden_state_size

    @property
    def is_blind(self):
        return False

    @property
    def num_recurrent_layers(self):
        return -1

    def _belief(self, observations, n, device):
        K.sync_pending()  # the belief vectors may still be in flight on the trainer's belief stream (ppo_trainer.py)
        belief = torch.zeros((n, self._hidden_size), device=device)
        if self._use_label_belief:
            cb = observations[CATEGORY_BELIEF]
            if self._normalize_category_distribution:
                cb = nn.functional.softmax(cb, dim=1)
            belief[:, :21] = cb
        if self._use_location_belief:
            belief[:, 21:23] = observations[LOCATION_BELIEF]
        return belief

    def forward(self, observations, rnn_hidden_states, prev_actions, masks, ext_memory, ext_memory_masks):
        x = self.get_features(observations, prev_actions)
        belief = self._belief(observations, x.shape[0], x.device)
        x_att = self.smt_state_encoder(x, ext_memory, ext_memory_masks, goal=belief)
        return x_att, rnn_hidden_states, x

    def pretrained_initialization(self, path):
        logging.info(f"AudioNavSMTNet ===> Loading pretrained model from {path}")
        state_dict = torch.load(path, map_location="cpu")["state_dict"]
        cleaned = {k[len("actor_critic.net."):]: v for k, v in state_dict.items() if "actor_critic.net." in k}
        self.load_state_dict(cleaned, strict=False)

    def freeze_encoders(self):
        """Freeze goal, visual and action encoders. Pose / fusion / transformer stay trainable (policy.py:643-653)."""
        for p in itertools.chain(self.goal_encoder.parameters(), self.visual_encoder.parameters(),
                                 self.action_encoder.parameters()):
            p.requires_grad = False

    def set_eval_encoders(self):
        self.goal_encoder.eval()
        self.visual_encoder.eval()

    def get_features(self, observations, prev_actions, extra_cols=0):
        """policy.py:660-674: [visual 128 | action 16 | audio 128 | (category 21) | pose 4] written straight into
        the column slices of one feature matrix (no torch.cat); ``extra_cols`` trailing columns are left for the
        caller (query-state embedding of the option net)."""
        n = observations[POSE].shape[0]
        dev = observations[POSE].device
        enc_trainable = torch.is_grad_enabled() and any(
            p.requires_grad for p in itertools.chain(self.goal_encoder.parameters(), self.visual_encoder.parameters(),
                                                     self.action_encoder.parameters()))
        if enc_trainable:
            # training encoders (savi_pretraining / savi_interactive yaml: freeze_encoders False): autograd path
            parts = [self.visual_encoder(observations)]
            if prev_actions.shape[1] == self._action_size:
                onehot = prev_actions.float()
            else:
                onehot = torch.zeros(n, self._action_size, device=dev).scatter_(1, prev_actions.long(), 1.0)
            parts.append(cuda_linear(onehot, self.action_encoder.weight, self.action_encoder.bias))
            parts.append(self.goal_encoder(observations))
            if self._use_category_input:
                parts.append(observations[CATEGORY])
            parts.append(observations[POSE])
            if extra_cols:
                parts.append(torch.zeros(n, extra_cols, device=dev))
            return torch.cat(parts, dim=1)
        with torch.no_grad():
            x = self._take_prefetch(observations, n, self._base_feature_size + extra_cols)
            if x is None:
                x = torch.empty((n, self._base_feature_size + extra_cols), device=dev, dtype=torch.float32)
                self._observation_features_into(x, observations)
            if prev_actions.shape[1] == self._action_size:  # already one-hot (policy.py:629-630)
                K.linear(prev_actions.float().contiguous(), self.action_encoder.weight, self.action_encoder.bias,
                         out=x[:, 128:144])
            else:
                K.onehot_linear(prev_actions, self.action_encoder.weight, self.action_encoder.bias, x[:, 128:144])
        return x

    def _observation_features_into(self, x, observations, visual=True, rest=True):
        """The columns of the feature row that depend on the observation only: visual 0:128, audio 144:272,
        (category), pose."""
        if visual:
            self.visual_encoder(observations, out=x[:, 0:128])
        if not rest:
            return
        self.goal_encoder(observations, out=x[:, 144:272])
        col = 272
        if self._use_category_input:
            K.copy_cols(observations[CATEGORY].contiguous(), x[:, col:col + 21])
            col += 21
        K.copy_cols(observations[POSE].contiguous(), x[:, col:col + 4])

    # ---- encoder prefetch: the observation-only columns of step s+1 are enqueued on a side stream as soon as the
    # environment has returned observation s+1 (trainer: right after ``envs.step``), so that the two visual ResNet-18s
    # and the audio CNN run next to the audio rendering / belief networks / storage insert instead of in front of the
    # next step's transformer.  Same kernels on the same data: the features are identical to the in-order path.
    # PPO.update uses the same mechanism one minibatch ahead (frozen encoders only).
    _PREFETCH_KEYS = ("rgb", "depth", SPECTROGRAM, POSE)

    def observation_key(self, observations):
        return tuple(observations[k].data_ptr() for k in self._PREFETCH_KEYS if k in observations)

    def _prefetched(self):
        d = self.__dict__.get("_prefetch")
        if d is None:
            d = self.__dict__["_prefetch"] = {}
        return d

    @torch.no_grad()
    def prefetch_observation_features(self, observations, key, stream, extra_cols=0, visual_event=None, between=None):
        """``observations``: what the environment returned (or the next minibatch of a PPO update); ``key``:
        ``observation_key`` of the tensors the consuming ``act`` / ``get_value`` / ``evaluate_actions`` call will be
        given (rollout: the storage slots these observations are copied into).  All prefetches must be enqueued on
        the same ``stream`` (the encoders' workspaces are per network).
        ``visual_event``: optional event after which the frames are complete (an environment that produces them on
        its own stream): the visual encoders then start without waiting for the rest of the current stream (the audio
        rendering); everything else waits for the current stream.  ``between``: optional callable run by the host
        after the visual encoders have been enqueued and before the remaining columns are (the trainer enqueues the
        belief networks there: they are the longer dependency chain of the next step)."""
        main = torch.cuda.current_stream()
        if visual_event is not None:
            stream.wait_event(visual_event)
        else:
            stream.wait_stream(main)
        n, dev = observations[POSE].shape[0], observations[POSE].device
        for k in self._PREFETCH_KEYS + (CATEGORY,):
            v = observations.get(k)
            if torch.is_tensor(v) and v.is_cuda:
                v.record_stream(stream)
        with torch.cuda.stream(stream):
            x = torch.empty((n, self._base_feature_size + extra_cols), device=dev, dtype=torch.float32)
            self._observation_features_into(x, observations, visual=True, rest=False)
        if between is not None:
            between()
        # the audio CNN and the pose / category columns only need what the main stream has produced (the spectrogram):
        # on a stream of their own they run NEXT TO the visual ResNet-18 chains instead of behind them (disjoint columns
        # of x); the prefetch is complete when both streams are
        rest = self.__dict__.get("_rest_stream")
        if rest is None:
            rest = self.__dict__["_rest_stream"] = torch.cuda.Stream()
        rest.wait_stream(main)
        x.record_stream(rest)
        for k in (SPECTROGRAM, POSE, CATEGORY):
            v = observations.get(k)
            if torch.is_tensor(v) and v.is_cuda:
                v.record_stream(rest)
        with torch.cuda.stream(rest):
            self._observation_features_into(x, observations, visual=False, rest=True)
            ev_rest = torch.cuda.Event()
            ev_rest.record(rest)
        with torch.cuda.stream(stream):
            stream.wait_event(ev_rest)
            ev = torch.cuda.Event()
            ev.record(stream)
        self._prefetched()[key] = (x, ev)

    def join_prefetch(self):
        """Make the current stream wait for every enqueued prefetch WITHOUT consuming it (a rollout step captured into
        a CUDA graph must have all its side-stream work joined before the capture ends; the next step's ``act`` then
        finds the features complete)."""
        cur = torch.cuda.current_stream()
        d = self._prefetched()
        for k, (x, ev) in list(d.items()):
            if ev is not None:
                cur.wait_event(ev)
                d[k] = (x, None)  # joined: the consumer must not wait on an event that belongs to an ended capture

    def drop_prefetch(self):
        d = self._prefetched()
        cur = torch.cuda.current_stream()
        for _x, ev in d.values():  # the encoders' workspaces are per network: whatever was enqueued must finish first
            if ev is not None:
                cur.wait_event(ev)
        d.clear()

    def _take_prefetch(self, observations, n, cols):
        d = self.__dict__.get("_prefetch")
        if not d:
            return None
        hit = d.pop(self.observation_key(observations), None)
        if hit is None or tuple(hit[0].shape) != (n, cols):
            # the caller is about to run the encoders itself: nothing enqueued earlier may still be using them
            if hit is not None and hit[1] is not None:
                torch.cuda.current_stream().wait_event(hit[1])
            self.drop_prefetch()
            return None
        x, ev = hit
        cur = torch.cuda.current_stream()
        if ev is not None:
            cur.wait_event(ev)
        x.record_stream(cur)
        return x


class AudioNavOptionNet(AudioNavSMTNet):
    """policy.py:919-1114 (pi_q): the SMT net whose current token carries the 32-d query-count embedding and whose
    stored memory rows carry the 32-d last-query embedding (memory_dim 308, fusion input 320).  The unused
    ``policy_selector`` / ``_qcnt_emb`` parameters are kept for checkpoint compatibility (Appendix A)."""

    def __init__(self, observation_space, action_space, hidden_size=128, query_count_emb_size=32, **kwargs):
        self._query_count_emb_size = query_count_emb_size
        kwargs.pop("use_query_count", None)
        super().__init__(observation_space, action_space, hidden_size=hidden_size, **kwargs)

    def _extra_feature_dims(self):
        return self._query_count_emb_size

    def _post_init(self, kwargs):
        self.policy_selector = nn.Linear(self._hidden_size, 2)
        self._qcnt_emb = nn.Embedding(2, self._query_count_emb_size)

    @property
    def qcnt_emb(self):
        return self._qcnt_emb

    def forward(self, observations, rnn_hidden_states, prev_actions, masks, ext_memory, ext_memory_masks, query_state,
                last_query_info):
        e = self._query_count_emb_size
        base = self._base_feature_size
        # the reference builds x_query = cat([x, query_state]) under torch.no_grad() (policy.py:1034-1036): pi_q never
        # back-propagates into its encoders, whatever freeze_encoders says -> the features are computed without a graph
        # (and therefore by the fused inference path), only the scene-memory transformer and the heads train
        with torch.no_grad():
            x = self.get_features(observations, prev_actions, extra_cols=e)  # [x (276) | query_state (32)]
            K.copy_cols(query_state.contiguous(), x[:, base:base + e])
        belief = self._belief(observations, x.shape[0], x.device)
        x_att = self.smt_state_encoder(x, ext_memory, ext_memory_masks, goal=belief)
        with torch.no_grad():  # memory rows: [x | last_query_info] (policy.py:1062-1063)
            x_for_memory = x.detach().clone()
            K.copy_cols(last_query_info.contiguous(), x_for_memory[:, base:base + e])
        return x_att, rnn_hidden_states, x_for_memory


class AudioNavDialogNet(AudioNavSMTNet):
    """policy.py:676-917 (pi_l): the SMT net followed by the CLIP-embedded dialog branch.  ``clip.*`` holds the text
    tower only (frozen, ddppo_trainer.py:401-403); ``dialog_layer`` maps its 512-d embedding to the hidden size and
    ``dialog_state_encoder`` attends over the K-step state memory."""

    def __init__(self, observation_space, action_space, hidden_size=128, num_steps=5, clip_layers=12, **kwargs):
        kwargs.pop("use_category_input", None)  # never appended by pi_l's get_features (policy.py:902-917)
        self._clip_layers = clip_layers
        self._num_steps = num_steps
        super().__init__(observation_space, action_space, hidden_size=hidden_size, **kwargs)

    def _post_init(self, kwargs):
        self.clip = CLIPTextTower(layers=self._clip_layers)
        self.dialog_layer = nn.Linear(512, self._hidden_size)
        self.dialog_state_encoder = DialogStateEncoder(self._hidden_size + self._hidden_size,
                                                       dim_feedforward=self._hidden_size, **kwargs)

    def encode_scene(self, observations, prev_actions, ext_memory, ext_memory_masks):
        """The part of pi_l that does not depend on the dialog: features, belief vector, scene-memory transformer.  The
        interactive trainer runs it on a side stream while pi_q decides whether a query fires."""
        x = self.get_features(observations, prev_actions)
        belief = self._belief(observations, x.shape[0], x.device)
        x_att = self.smt_state_encoder(x, ext_memory, ext_memory_masks, goal=belief)
        return x, x_att, belief

    def forward(self, observations, rnn_hidden_states, prev_actions, masks, ext_memory, ext_memory_dialog,
                ext_memory_masks, all_dialog, agent_step, scene=None):
        if scene is None:
            scene = self.encode_scene(observations, prev_actions, ext_memory, ext_memory_masks)
        x, x_att, belief = scene
        if all_dialog is not None:
            if torch.is_grad_enabled():
                dialog_emb = self.clip.encode_text(all_dialog)  # no_grad, fp32 (policy.py:847-849)
            else:  # rollout: per-env embedding cache, only rows whose dialog changed are encoded
                dialog_emb = self.clip.encode_text_cached(all_dialog)
            dialog_emb = cuda_linear(dialog_emb, self.dialog_layer.weight, self.dialog_layer.bias)
        else:
            dialog_emb = None
        # policy.py:862 hands the SAME mask to the scene memory and to the dialog memory; the dialog memory may be
        # shorter than the mask (K slots): its slots are the first K mask columns
        Kd = (ext_memory_dialog.memory if isinstance(ext_memory_dialog, IndexedMemory) else ext_memory_dialog).shape[0]
        dmask = ext_memory_masks if ext_memory_masks.shape[1] == Kd else ext_memory_masks[:, :Kd]
        x_att_dialog = self.dialog_state_encoder(x_att, ext_memory_dialog, dmask, dialog_emb, agent_step, goal=belief)
        return x_att_dialog, rnn_hidden_states, x, x_att_dialog


class AudioNavSMTPolicy(Policy):
    def __init__(self, observation_space, action_space, hidden_size=128, **kwargs):
        super().__init__(AudioNavSMTNet(observation_space, action_space, hidden_size=hidden_size, **kwargs),
                         action_space.n)


class AudioNavDialogPolicy(Policy):
    """policy.py:334-344 (pi_l)."""

    def __init__(self, observation_space, action_space, hidden_size=128, **kwargs):
        super().__init__(AudioNavDialogNet(observation_space, action_space, hidden_size=hidden_size, **kwargs),
                         action_space.n)


class AudioNavOptionPolicy(Policy):
    """policy.py:346-356: all three action heads have 2 outputs."""

    def __init__(self, observation_space, action_space, hidden_size=128, **kwargs):
        super().__init__(AudioNavOptionNet(observation_space, action_space, hidden_size=hidden_size, **kwargs), 2)
