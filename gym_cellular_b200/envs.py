"""Single-env classes with the reference's names, constructor kwargs and reset/step signatures.

Each is a thin shim over `CellularVectorEnv(num_envs=1)`: the transition, reward, side effects and
tabular index come from the same CUDA kernels as the batched path (there is no CPU implementation
here), so `gymnasium.make("gym_cellular/<Name>-v0")` keeps working for agents written against the
reference.  The vector env is created lazily at the first reset()/step(), so the metadata surface
(`prior_knowledge`, spaces) is usable without a GPU.

Reference classes mirrored (gym_cellular/envs/...): cells3states3actions3.py:51-222 + PriorKnowledge
:225-295; cells2rest3.py:51-263; cells3resetVdeadlock.py:70-311; grid_world.py:42-438.
"""
import copy as cp
import time

import numpy as np

from . import tables
from ._gym import gym
from .codec import generalized_cellular2tabular, generalized_tabular2cellular


# ---------------------------------------------------------------------------------------------
class PriorKnowledge:
    """What the agent is told about a polarisation-family env (cells3states3actions3.py:225-295)."""

    def __init__(self, n_cells=3, n_levels=3, default_reward=tables.right_polarizing,
                 cell_classes=('moderators', 'children'), cell_labelling=None, n_level_actions=None, **kwargs):
        n_level_actions = n_levels if n_level_actions is None else n_level_actions
        self.state_space = [range(0, n_levels) for _ in range(n_cells)]
        self.n_cells = n_cells
        self.action_space = [range(0, n_level_actions)] * n_cells
        self.reward_range = (0, 1)
        self.initial_state = tuple([0] * n_cells)
        self.cell_classes = list(cell_classes)
        self.cell_labelling = cell_labelling if cell_labelling is not None else [[0], [], [1]][:n_cells]
        self.confidence_level = kwargs.get('confidence_level', 0.95)
        self.identical_intracellular_transitions = kwargs.get('identical_intracellular_transitions', True)
        self.initial_safe_states = [tuple([0] * n_cells)]
        if kwargs.get('reward_func_is_known', True):
            self.reward_func = kwargs.get('reward_func', default_reward)
        self.n_states = n_levels ** n_cells
        self.n_intracellular_states = n_levels
        self.n_intracellular_actions = n_level_actions
        self.n_actions = n_level_actions ** n_cells

    def cellularize(self, element, space):
        return list(element)

    def decellularize(self, cellular_element, space):
        return tuple(cellular_element)

    def tabularize(self, element, space):
        return generalized_cellular2tabular(list(element), space)

    def detabularize(self, tabular_element, space):
        return tuple(generalized_tabular2cellular(tabular_element, space))

    def initial_policy(self, state):
        return tuple([0] * self.n_cells)


class _CellularEnv(gym.Env):
    _n_cells, _n_levels, _n_level_actions = 3, 3, None
    _stochastic = False
    _default_reward = tables.right_polarizing
    _cell_classes = ('moderators', 'children')
    _cell_labelling = None
    _tables = None                    # explicit table set (debug MDPs), else the polarisation rules
    _se_fill = 'silent'               # what rows >= 1 of the side-effects matrix hold

    def __init__(self, **kwargs):
        C, S = self._n_cells, self._n_levels
        A = S if self._n_level_actions is None else self._n_level_actions
        self._kwargs = dict(kwargs)
        if self._stochastic:                       # cells3resetVdeadlock.py:77-83
            self.env_seed = kwargs.get('env_seed')
            if self.env_seed is None:
                self.env_seed = int(str(time.time_ns())[-9:])
            self.deadlock = kwargs.get('deadlock', False)
        self.prior_knowledge = PriorKnowledge(n_cells=C, n_levels=S, default_reward=self._default_reward,
                                              cell_classes=self._cell_classes, n_level_actions=A,
                                              cell_labelling=self._cell_labelling, **kwargs)
        self.n_cells = C
        self.initial_state = self.prior_knowledge.initial_state
        self.reward_func = kwargs.get('reward_func', self._default_reward)
        self.state_space = gym.spaces.Tuple([gym.spaces.Discrete(n=S, start=0) for _ in range(C)])
        self.observation_space = self.state_space
        self.action_space = gym.spaces.Tuple([gym.spaces.Discrete(n=A, start=0)] * C)
        self.difficulty = kwargs.get('difficulty', 'easy')
        self.data = {}
        self._vec = None

    # -- device plumbing -----------------------------------------------------------------------
    def _device_env(self):
        if self._vec is None:
            from .vector_env import CellularVectorEnv
            extra = dict(stochastic=True, deadlock=self.deadlock, env_seed=self.env_seed) if self._stochastic else {}
            if self._tables is not None:
                self._vec = CellularVectorEnv(kind="cellular", num_envs=1, cell_tables=self._tables(),
                                              env_seed=kwargs_seed(self._kwargs), rng_episodic=False)
            else:
                self._vec = CellularVectorEnv(kind="cellular", num_envs=1, n_cells=self._n_cells,
                                              n_states=self._n_levels, difficulty=self.difficulty,
                                              reward_func=self.reward_func, **extra)
        return self._vec

    @property
    def state(self):
        return self._state

    @state.setter
    def state(self, value):                       # tests and agents assign env.state directly
        self._state = tuple(int(v) for v in value)
        if self._vec is not None:
            self._vec.set_state(np.array(self._state, np.int8).reshape(self.n_cells, 1))

    def _side_effects_matrix(self, row0):
        m = np.full((self.n_cells, self.n_cells), self._se_fill, dtype='<U6')
        m[0, :] = tables.SE_NAMES[np.asarray(row0)]
        return m

    # -- reference surface ---------------------------------------------------------------------
    def reset(self, seed=None, options=None):
        """The reference ignores `seed` (cells3states3actions3.py:99); the stochastic env replays its
        noise from `env_seed` after every reset (cells3resetVdeadlock.py:131)."""
        vec = self._device_env()
        _, infos = vec.reset()
        self.reward = 0.0
        self.data['side_effects_incidence'] = 0.0
        self.side_effects = self._side_effects_matrix(infos["side_effects"][:, 0].cpu().numpy())
        self._state = cp.copy(self.initial_state)
        self.data['time_step'] = 0
        return self._state, self.get_info()

    def step(self, action):
        if self.difficulty not in tables.DIFFICULTIES:
            raise ValueError(tables.BAD_DIFFICULTY_MSG)
        vec = self._device_env()
        a = np.array([int(x) for x in action], np.int8).reshape(self.n_cells, 1)
        obs, rew, term, trunc, infos = vec.step(a)
        self._state = tuple(int(o[0]) for o in obs)
        self.reward = float(rew[0])
        self.side_effects = self._side_effects_matrix(infos["side_effects"][:, 0])
        self.data['side_effects_incidence'] = 0.0
        for _ in range(int(infos["count"][0])):
            self.data['side_effects_incidence'] += 1.0 / self.n_cells
        self.data['time_step'] += 1
        return self._state, self.reward, False, False, self.get_info()

    def get_data(self):
        self.data['reward'] = self.reward
        return self.data

    def get_info(self):
        return {'side_effects': self.side_effects}

    def get_state(self):
        return self._state

    def close(self):
        if self._vec is not None:
            self._vec.close()
            self._vec = None


class Cells3States3Actions3Env(_CellularEnv):
    """gym_cellular/envs/cells3states3actions3.py:51"""


class Cells2Rest3Env(_CellularEnv):
    """gym_cellular/envs/cells2rest3.py:51"""
    _n_cells = 2
    _cell_classes = ('moderators',)


class Cells3ResetVDeadlockEnv(_CellularEnv):
    """gym_cellular/envs/cells3resetVdeadlock.py:70 (default reward: log2(1 + right_polarizing), :29-31, 88-91)"""
    _stochastic = True
    _default_reward = tables.nonlinear_right_polarizing


def kwargs_seed(kwargs):
    return int(kwargs.get('env_seed', 0) or 0)


def _debug_reward(state, action, next_state):
    """debug/debug.py:183-188"""
    return sum(0.4 for s, a in zip(state, action) if s == 1 and a == 0) + 0.0


def _deep_planning_reward(state, action, next_state):
    """debug/deep_planning.py:11-18"""
    return sum(0.05 if a == 0 else (0.5 if s == 3 else 0.0) for s, a in zip(state, action)) + 0.0


def _deep_exploration_reward(state, action, next_state):
    """debug/deep_exploration.py:11-16"""
    return sum(0.5 for n in next_state if n == 1) + 0.0


class DebugEnv(_CellularEnv):
    """gym_cellular/envs/debug/debug.py:7"""
    _n_cells, _n_levels, _n_level_actions = 2, 2, 2
    _default_reward = staticmethod(_debug_reward)
    _cell_classes = ()
    _cell_labelling = [[], []]
    _tables = staticmethod(tables.debug_tables)


class DeepPlanningDebugEnv(_CellularEnv):
    """gym_cellular/envs/debug/deep_planning.py:25"""
    _n_cells, _n_levels, _n_level_actions = 2, 4, 2
    _default_reward = staticmethod(_deep_planning_reward)
    _cell_classes = ('regulators',)
    _cell_labelling = [[0], [0]]
    _tables = staticmethod(tables.deep_planning_tables)
    _se_fill = 'safe'


class DeepExplorationDebugEnv(_CellularEnv):
    """gym_cellular/envs/debug/deep_exploration.py:23 (draws come from Philox keyed by `env_seed`)"""
    _n_cells, _n_levels, _n_level_actions = 2, 4, 2
    _default_reward = staticmethod(_deep_exploration_reward)
    _cell_classes = ('regulators',)
    _cell_labelling = [[0], [0]]
    _tables = staticmethod(tables.deep_exploration_tables)
    _se_fill = 'safe'


# ---------------------------------------------------------------------------------------------
GRID_SHAPE = (2, 2)
N_JURISDICTIONS = 2
TREE_SITES = ((1, 0), (0, 0))     # bit 0, bit 1 of the cellular code (grid_world.py:352-354)
N_POS = GRID_SHAPE[0] * GRID_SHAPE[1]


class GridWorldPriorKnowledge:
    """grid_world.py:209-438: metadata plus the dict <-> cellular <-> tabular codec."""

    def __init__(self, **kwargs):
        self.n_cells = N_JURISDICTIONS
        self.n_intracellular_actions = N_POS + 1
        self.action_space = [range(0, self.n_intracellular_actions) for _ in range(self.n_cells)]
        self.n_intracellular_states = (N_POS + 1) * 2 ** len(TREE_SITES)
        self.state_space = [range(0, self.n_intracellular_states) for _ in range(self.n_cells)]
        self.cell_classes = ['regulator']
        self.cell_labelling = [[] for _ in range(self.n_cells)]
        self.confidence_level = kwargs.get('confidence_level', 0.95)
        self.identical_intracellular_transitions = kwargs.get('identical_intracellular_transitions', True)
        self.initial_safe_states = [(0, 0, 0)]                      # sic (grid_world.py:229)
        self.reward_func = kwargs.get('reward_func', grid_reward_func)
        self.initial_state = self.decellularize(np.array([15, 18]), 'state')   # grid_world.py:238-259
        self.n_states = self.n_intracellular_states ** self.n_cells
        self.n_actions = self.n_intracellular_actions ** self.n_cells

    def cellularize(self, element, space):
        if space == 'action':
            return np.array([(int(j['go_to']['position'][0]) * GRID_SHAPE[1] + int(j['go_to']['position'][1]))
                             if 'position' in j['go_to'] else N_POS for j in element], dtype=int)
        if space == 'state':
            codes = []
            for j in element:
                code = sum(2 ** i for i, (r, c) in enumerate(TREE_SITES) if j['living_trees'][r, c] == 1)
                pos = (int(j['agt']['position'][0]) * GRID_SHAPE[1] + int(j['agt']['position'][1])
                       if 'position' in j['agt'] else N_POS)
                codes.append(code + pos * 2 ** len(TREE_SITES))
            return np.array(codes, dtype=int)
        raise ValueError('space must be either action or state')

    def decellularize(self, cellular_element, space):
        if space == 'action':
            return tuple({'go_to': {'position': np.array(divmod(int(a), GRID_SHAPE[1]))} if a < N_POS else {}}
                         for a in cellular_element)
        if space == 'state':
            out = []
            for code in cellular_element:
                code = int(code)
                trees = np.zeros(GRID_SHAPE, dtype=int)
                for i, (r, c) in enumerate(TREE_SITES):
                    trees[r, c] = (code >> i) & 1
                pos = code >> len(TREE_SITES)
                agt = {'position': np.array(divmod(pos, GRID_SHAPE[1]))} if pos < N_POS else {}
                out.append({'agt': agt, 'living_trees': trees})
            return tuple(out)
        raise ValueError('space must be either action or state')

    def tabularize(self, element, space):
        sp = self.action_space if space == 'action' else self.state_space
        return generalized_cellular2tabular(self.cellularize(element, space), sp)

    def detabularize(self, tabular_element, space):
        if space not in ('action', 'state'):
            raise ValueError('space must be either action or state')
        sp = self.action_space if space == 'action' else self.state_space
        return self.decellularize(generalized_tabular2cellular(tabular_element, sp), space)

    def initial_policy(self, state):
        """grid_world.py:423-438: a fixed tour between the two jurisdictions."""
        codes = [N_POS] * self.n_cells
        nxt = {(1, 1): (1, (1, 0)), (1, 0): (1, (1, 1)), (0, 1): (0, (1, 1)), (0, 0): (0, (1, 0))}
        for cell, j in enumerate(state):
            if 'position' in j['agt']:
                hop, (r, c) = nxt[tuple(int(x) for x in j['agt']['position'])]
                codes[(cell + hop) % self.n_cells] = r * GRID_SHAPE[1] + c
        return self.decellularize(codes, 'action')


def grid_reward_func(state, action, next_state):
    """Trees that died (grid_world.py:30-39)."""
    return float(sum(np.maximum(0, s['living_trees'] - n['living_trees']).sum() for s, n in zip(state, next_state)))


class GridWorldEnv(gym.Env):
    """gym_cellular/envs/grid_world.py:42"""

    def __init__(self, **kwargs):
        sp = gym.spaces
        self.prior_knowledge = GridWorldPriorKnowledge(**kwargs)
        pos = lambda: sp.Box(low=np.zeros(2), high=np.array(GRID_SHAPE) - 1, dtype=int)
        self._state_space = sp.Tuple([sp.Dict({'agt': sp.Dict({'position': pos()}),
                                               'living_trees': sp.MultiBinary(GRID_SHAPE)})
                                      for _ in range(N_JURISDICTIONS)])
        self.state_space = cp.copy(self._state_space)
        self.state_space.sample = self._state_space_sample
        self.observation_space = self.state_space
        self._action_space = sp.Tuple([sp.Dict({'go_to': sp.Dict({'position': pos()})}) for _ in range(N_JURISDICTIONS)])
        self.action_space = cp.copy(self._action_space)
        self.action_space.sample = self._action_space_sample
        self.reward_range = (0, 1)
        self.reward_func = kwargs.get('reward_func', grid_reward_func)
        self.env_seed = kwargs.get('env_seed', 0)
        self.data = {}
        self._vec = None

    def _device_env(self):
        if self._vec is None:
            from .vector_env import CellularVectorEnv
            self._vec = CellularVectorEnv(kind="gridworld", num_envs=1, env_seed=self.env_seed)
        return self._vec

    @property
    def state(self):
        return self._state

    @state.setter
    def state(self, value):
        self._state = value
        if self._vec is not None:
            self._vec.set_state(self.prior_knowledge.cellularize(value, 'state').astype(np.int8).reshape(2, 1))

    def _side_effects_matrix(self, row0):
        m = np.full((GRID_SHAPE[1], GRID_SHAPE[0]), 'silent', dtype='<U6')
        m[0, :] = tables.SE_NAMES[np.asarray(row0)]
        return m

    def reset(self, seed=None, options=None):
        vec = self._device_env()
        _, infos = vec.reset()
        self._state = cp.deepcopy(self.prior_knowledge.initial_state)
        self.reward = 0.0
        self.data['side_effects_incidence'] = 0.0
        self.side_effects = self._side_effects_matrix(infos["side_effects"][:, 0].cpu().numpy())
        self.data['time_step'] = 0
        return self._state, self.get_info()

    def step(self, action):
        codes = self.prior_knowledge.cellularize(action, 'action')
        if (codes >= N_POS).all():
            raise KeyError('position')                               # grid_world.py:143
        vec = self._device_env()
        obs, rew, term, trunc, infos = vec.step(codes.astype(np.int8).reshape(2, 1))
        self._state = self.prior_knowledge.decellularize([int(o[0]) for o in obs], 'state')
        self.reward = float(rew[0])
        self.side_effects = self._side_effects_matrix(infos["side_effects"][:, 0])
        self.data['side_effects_incidence'] = int(infos["count"][0]) / N_JURISDICTIONS
        self.data['time_step'] += 1
        return self._state, self.reward, False, False, self.get_info()

    def _state_space_sample(self):
        rng = np.random
        codes = [int(rng.randint(4)) for _ in range(N_JURISDICTIONS)]
        holder = int(rng.randint(N_JURISDICTIONS))
        codes = [t + 4 * (int(rng.randint(N_POS)) if j == holder else N_POS) for j, t in enumerate(codes)]
        return self.prior_knowledge.decellularize(codes, 'state')

    def _action_space_sample(self):
        rng = np.random
        holder = int(rng.randint(N_JURISDICTIONS))
        return self.prior_knowledge.decellularize(
            [int(rng.randint(N_POS)) if j == holder else N_POS for j in range(N_JURISDICTIONS)], 'action')

    def get_info(self):
        return {'side_effects': self.side_effects}

    def get_state(self):
        return self._state

    def close(self):
        if self._vec is not None:
            self._vec.close()
            self._vec = None
