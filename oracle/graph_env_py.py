"""CPU restatement of the SoundSpaces graph walk, oracle actions and reward — TEST INFRASTRUCTURE.

Follows, one env at a time like the reference:
  soundspaces/simulator.py:496-517   ``step``: FORWARD to the neighbour whose direction equals the orientation, turns
  soundspaces/simulator.py:594-603   ``get_orientation`` / ``azimuth_angle``
  soundspaces/simulator.py:758-787   ``compute_oracle_actions`` on ``nx.shortest_path``
  ss_baselines/common/environments.py:98-135  ``get_reward`` (slack, geodesic progress, success, query penalties)
on a ``networkx`` graph whose nodes carry ``point = (x, y, z)`` exactly as the reference's scene graphs do."""
from __future__ import annotations

import math

import numpy as np

STOP, FORWARD, LEFT, RIGHT = 0, 1, 2, 3


def build_nx_graph(points, nbr):
    import networkx as nx
    g = nx.Graph()
    for i, (x, z) in enumerate(points):
        g.add_node(i, point=(float(x), 0.0, float(z)))
    for i in range(len(points)):
        for d in range(4):
            j = int(nbr[i, d])
            if j >= 0:
                g.add_edge(i, j)
    return g


def direction_deg(graph, a, b):  # simulator.py:506 / :768
    p1, p2 = graph.nodes[a]["point"], graph.nodes[b]["point"]
    return int(np.around(np.rad2deg(np.arctan2(p2[2] - p1[2], p2[0] - p1[0])))) % 360


class RefGraphEnv:
    def __init__(self, graph, grid_size, reward_cfg, max_steps=500):
        self.graph, self.grid_size, self.cfg, self.max_steps = graph, grid_size, reward_cfg, max_steps

    def reset(self, start, rotation_angle, source):
        self.node, self.rotation_angle, self.source = start, rotation_angle % 360, source
        self.step_count = 0
        self.stop_called = False
        self.prev_dist = self.distance()  # environments.py:66-68
        self.is_queried, self.query_num, self.cons_reward = False, 0, 0.0

    def get_orientation(self):  # :594-596
        return (270 - self.rotation_angle) % 360

    def azimuth_angle(self):  # :598-603
        return -(self.rotation_angle + 0) % 360

    def distance(self):  # :736-745
        import networkx as nx
        return nx.shortest_path_length(self.graph, self.node, self.source) * self.grid_size

    def step(self, action):  # :496-517
        if action == STOP:
            self.stop_called = True
        elif action == FORWARD:
            for neighbor in self.graph[self.node]:
                if direction_deg(self.graph, self.node, neighbor) == self.get_orientation():
                    self.node = neighbor
                    break
        elif action == LEFT:
            self.rotation_angle = (self.rotation_angle + 90) % 360
        elif action == RIGHT:
            self.rotation_angle = (self.rotation_angle - 90) % 360
        self.step_count += 1

    def get_reward(self):  # environments.py:98-135
        c = self.cfg
        reward = 0.0
        if c["WITH_TIME_PENALTY"]:
            reward += c["SLACK_REWARD"]
        if c["WITH_DISTANCE_REWARD"]:
            cur = self.distance()
            reward += (self.prev_dist - cur) * c["DISTANCE_REWARD_SCALE"]
            self.prev_dist = cur
        if self.stop_called and self.node == self.source:
            reward += c["SUCCESS_REWARD"]
        if c["WITH_QUERY_CONSTRAINT"] and self.is_queried:
            if self.query_num <= c["NUM_TOTAL_QUERY"]:
                if c["SOFT_QUERY_REWARD"]:
                    reward += (self.query_num / c["NUM_TOTAL_QUERY"]) * (math.exp(-c["NUM_TOTAL_QUERY"]) + c["QUERY_REWARD"])
            else:
                reward += math.exp(-self.query_num) + c["QUERY_REWARD"]
            if c["CONSECUTIVE_CONSTRAINT_REWARD"]:
                reward += self.cons_reward
        return reward

    def done(self):
        return self.stop_called or self.step_count >= self.max_steps

    def oracle_actions_along(self, path):  # :758-787 for a given node path
        actions, orientation = [], self.get_orientation()
        for a, b in zip(path[:-1], path[1:]):
            direction = direction_deg(self.graph, a, b)
            if direction == orientation:
                pass
            elif (direction - orientation) % 360 == 270:
                orientation = (orientation - 90) % 360
                actions.append(LEFT)
            elif (direction - orientation) % 360 == 90:
                orientation = (orientation + 90) % 360
                actions.append(RIGHT)
            elif (direction - orientation) % 360 == 180:
                orientation = (orientation - 180) % 360
                actions.append(RIGHT)
                actions.append(RIGHT)
            actions.append(FORWARD)
        actions.append(STOP)
        return actions

    def compute_oracle_actions(self):
        import networkx as nx
        return self.oracle_actions_along(nx.shortest_path(self.graph, source=self.node, target=self.source))
