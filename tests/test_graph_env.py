"""Graph-walk environment: scene tables against networkx (CPU), the device step kernel against
the restatement of soundspaces/simulator.py:496-517,:758-787 + ss_baselines/common/environments.py:98-135 (GPU)."""
import numpy as np
import pytest
import torch

from oracle import graph_env_py as G


def _scene(seed=3, grid=6):
    from avlen_b200.graph_env import GraphScene
    return GraphScene(grid=grid, grid_size=1.0, obstacle_frac=0.2, seed=seed)


def test_scene_tables_match_networkx():
    import networkx as nx
    sc = _scene()
    g = G.build_nx_graph(sc.points, sc.nbr)
    assert nx.is_connected(g) and sc.V > 10
    sp = dict(nx.all_pairs_shortest_path_length(g))
    for a in range(sc.V):
        for b in range(sc.V):
            assert sc.hops[a, b] == sp[a][b]
            if a != b:
                d = int(sc.next_dir[b, a])  # first edge from a towards target b
                u = sc.nbr[a, d]
                assert u >= 0 and sp[u][b] == sp[a][b] - 1
                assert G.direction_deg(g, a, int(u)) == 90 * d  # the table's direction convention is simulator.py:506's
    assert (np.diag(sc.next_dir) == -1).all()


def test_table_oracle_plan_is_a_reference_oracle_plan():
    """Walking the product's first-edge table gives a node path of minimal length; the reference's own action rule
    (:768-784) on that path reaches the goal and ends with STOP.  Where the shortest path is unique the plan equals
    ``compute_oracle_actions`` on nx.shortest_path."""
    import networkx as nx
    sc = _scene(seed=5)
    g = G.build_nx_graph(sc.points, sc.nbr)
    rng = np.random.default_rng(0)
    env = G.RefGraphEnv(g, 1.0, {})
    for _ in range(40):
        a, b, rot = int(rng.integers(sc.V)), int(rng.integers(sc.V)), int(rng.integers(4)) * 90
        path = [a]
        while path[-1] != b:
            path.append(int(sc.nbr[path[-1], sc.next_dir[b, path[-1]]]))
        assert len(path) - 1 == sc.hops[b, a]
        env.reset(a, rot, b)
        plan = env.oracle_actions_along(path)
        assert plan[-1] == G.STOP and plan.count(G.FORWARD) == sc.hops[b, a]
        if len(list(nx.all_shortest_paths(g, a, b))) == 1:
            assert plan == env.compute_oracle_actions()


@pytest.mark.gpu
def test_graph_env_step_kernel_matches_reference_restatement():
    from avlen_b200.graph_env import REWARD_DEFAULTS, GraphVectorEnv
    n, steps = 8, 120
    sc = _scene(seed=7, grid=5)
    env = GraphVectorEnv(n, "cuda", scene=sc, episodes_per_env=16, max_episode_steps=25, seed=11)
    env.reset()
    g = G.build_nx_graph(sc.points, sc.nbr)
    cfg = dict(REWARD_DEFAULTS)
    refs = [G.RefGraphEnv(g, 1.0, cfg, max_steps=25) for _ in range(n)]
    ep_start, ep_rot, ep_src = env._ep_start.cpu().numpy(), env._ep_rot.cpu().numpy(), env._ep_source.cpu().numpy()
    cursor = [0] * n
    for i, r in enumerate(refs):
        r.reset(int(ep_start[i, 0]), 90 * int(ep_rot[i, 0] & 3), int(ep_src[i, 0]))
    rng = np.random.default_rng(1)
    oracle0 = env.compute_oracle_actions().cpu().numpy()
    for t in range(steps):
        # two thirds of the actions follow the oracle (so that episodes succeed), the rest is random
        o_now = env.compute_oracle_actions().cpu().numpy()
        acts = np.where(rng.random(n) < 0.66, o_now, rng.integers(0, 4, n))
        isq = rng.random(n) < 0.3
        qn = rng.integers(0, 7, n)
        cons = np.where(rng.random(n) < 0.5, -0.25, 0.0).astype(np.float32)
        env.set_is_queried(torch.from_numpy(isq).cuda())
        env.set_query_num(torch.from_numpy(qn).cuda())
        env.set_constraint_reward(torch.from_numpy(cons).cuda())
        obs, rew, dones = env.step(torch.from_numpy(acts).cuda().view(n, 1))
        rew, dones = rew.cpu().numpy().reshape(-1), dones.cpu().numpy()
        node, rot = env._gstate[0].cpu().numpy(), env._gstate[1].cpu().numpy()
        oracle = env.compute_oracle_actions().cpu().numpy()
        td = env.target_distance().cpu().numpy()
        for i, r in enumerate(refs):
            r.is_queried, r.query_num, r.cons_reward = bool(isq[i]), int(qn[i]), float(cons[i])
            r.step(int(acts[i]))
            want = r.get_reward()
            assert abs(rew[i] - want) < 1e-5, (t, i, rew[i], want)
            assert bool(dones[i]) == r.done(), (t, i)
            if r.done():  # auto-reset
                cursor[i] = (cursor[i] + 1) % ep_start.shape[1]
                c = cursor[i]
                r.reset(int(ep_start[i, c]), 90 * int(ep_rot[i, c] & 3), int(ep_src[i, c]))
            assert node[i] == r.node and 90 * int(rot[i]) == r.rotation_angle, (t, i)
            assert abs(td[i] - r.distance()) < 1e-6
            # first oracle action: the reference's rule (:768-787) applied to the table's next node
            if r.node == r.source:
                assert oracle[i] == G.STOP
            else:
                nxt = int(sc.nbr[r.node, sc.next_dir[r.source, r.node]])
                assert oracle[i] == r.oracle_actions_along([r.node, nxt])[0], (t, i)
            assert int(env._azimuth[i]) * 90 == r.azimuth_angle()
        assert np.array_equal(obs["pose"][:, 3].cpu().numpy(), env._gstate[3].cpu().numpy().astype(np.float32))
    assert dones is not None and oracle0.shape == (n,)
    env.close()
