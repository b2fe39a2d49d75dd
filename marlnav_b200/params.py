"""Parameter dicts for the drop-in ``Env``.

``Env`` consumes the reference's own ``params['env']`` dict unchanged
(/root/reference/marlnav/utils.py:257-282), so a user of the reference keeps
calling ``set_params(args)``.  This module only restates the reference's CLI
defaults (__main__.py:49-133) and its three built-in scenarios
(``-sn -1/0/1``: utils.py:17-115) so that tests, the benchmark and the smoke run
can build the same dicts on a machine where the reference is not installed.
"""
import copy
import math

# Geometry constants the reference hard-codes in Env.__init__ (environment.py:56-68).
GEOMETRY = dict(ob_risk_dist=60., ag_risk_dist=15., ob_coll_dist=50., ag_coll_dist=5.,
                agents_min_d=30., agents_max_d=50., max_at_prop_d=2., max_angle_diff=math.pi / 8,
                target_radius=30., cap_distance=0.1, bond_sharpness=1., ideal_dist=40.,
                init_dist=1200.)

# utils.py:17-33
TRIANGLE_INIT = dict(init_method='triangle', ags_cent_x=150., ags_cent_y=375., ags_dist=40.,
                     init_speed=3., tar_pos_x=1350., tar_pos_y=375., noisy_ags=False, ags_std=0.01,
                     angle_range=math.pi / 6, obst_min_x=500., obst_max_x=1000., obst_min_y=250.,
                     obst_max_y=500.)

# utils.py:35-62 (two identical envs; one obstacle each)
MOCK_INIT_0 = dict(
    init_method='mock_init',
    mock_states=[[[550., 100., 0., 1., 0.], [750., 100., 0., 1., 0.], [950., 100., 0., 1., 5.]]] * 2,
    mock_obstacles=[[[1400., 375.]]] * 2,
    mock_target=[[[1400., 700.]]] * 2)

# utils.py:64-91
_S3 = math.sqrt(3)
_V = 2 * 300. * math.sin(math.radians(0.9))
MOCK_INIT_1 = dict(
    init_method='mock_init',
    mock_states=[
        [[750. - 300. / _S3, 375., 0., 1., 3. / math.sin(math.pi / 3)],
         [750., 375., 0., 1., 3.],
         [750. + 300. / _S3, 375., 0., 1., 3. / math.sin(math.pi / 3)]],
        [[450, 675., 1., 0., _V], [750., 675., 0., -1., 6.], [1050., 675., -1., 0., _V]]],
    mock_obstacles=[[[900., 475.]], [[750., 75.]]],
    mock_target=[[[750., 675.]], [[750., 475.]]])

# utils.py:93-115
CONST_SAMPLER = dict(sample_method='const_sampler')
MOCK_SAMPLER_0 = dict(sampler_num=0, sample_method='mock_sampler',
                      actions=[[[0., 5.], [0., 0.1], [0., -0.05]],
                               [[0., 5.], [0., 0.1], [0., -100.]]])
MOCK_SAMPLER_1 = dict(sampler_num=1, sample_method='mock_sampler',
                      actions=[[[0., 0.], [0., 0.], [0., 0.]],
                               [[-math.radians(1.8), 0.], [0., 0.], [math.radians(1.8), 0.]]])


def default_env_params(num_parallel=2, num_agents=3, num_obstacles=3, sampler_num=-1,
                       sampling_style='sampler', device='cuda', **overrides):
    """The dict ``set_env_params`` would build from the CLI defaults for the given
    ``-np/-na/-no/-sn/-sa`` (utils.py:217-282).  For ``sampler_num`` 0/1 the
    reference scenarios fix ``num_parallel=2, num_agents=3, num_obstacles=1``."""
    if sampler_num in (0, 1):
        num_parallel, num_agents, num_obstacles = 2, 3, 1
    env = dict(device=device, num_parallel=num_parallel, num_agents=num_agents,
               num_obstacles=num_obstacles, x_bound=1500.0, y_bound=750.0, max_step=1000,
               episode_len=200, min_speed=3., max_speed=10., min_accel=-0.5, max_accel=0.5,
               risk_factor=0., distance_factor=0., heading_factor=500., target_factor=500.,
               soft_factor=500., bond_factor=10.)
    if sampler_num == -1:
        init = dict(TRIANGLE_INIT, num_parallel=num_parallel, num_obs=num_obstacles)
        sampler = None if sampling_style == 'policy' else dict(
            CONST_SAMPLER, num_parallel=num_parallel, num_agents=num_agents)
    elif sampler_num == 0:
        init, sampler = copy.deepcopy(MOCK_INIT_0), dict(copy.deepcopy(MOCK_SAMPLER_0), max_step=1000)
    elif sampler_num == 1:
        init, sampler = copy.deepcopy(MOCK_INIT_1), dict(copy.deepcopy(MOCK_SAMPLER_1), max_step=1000)
    else:
        raise ValueError(sampler_num)
    init['device'] = device
    if sampler is not None:
        sampler['device'] = device
    env['init'], env['sampler'] = init, sampler
    env.update(overrides)
    return env


def ring_template(num_agents, cx=150., cy=375., speed=3., spacing=40.):
    """(A,5) reset template for teams the reference's 3-agent triangle cannot express
    (utils.py:350-368 is hard-wired to 3): agents on a ring, neighbours ``spacing``
    apart, heading (1,0).  Used for BASELINE.json's scaled scene (8 agents)."""
    r = (spacing / 2.0) / math.sin(math.pi / num_agents)
    return [[cx + r * math.cos(2.0 * math.pi * i / num_agents),
             cy + r * math.sin(2.0 * math.pi * i / num_agents), 1.0, 0.0, speed]
            for i in range(num_agents)]


def template_env_params(num_parallel, num_agents, num_obstacles, agent_template=None, device='cuda',
                        **overrides):
    """Env params with an explicit (A,5) agent reset template (``init_method='template'``,
    an extension: same obstacle box / target / Philox resets as the triangle scenario)."""
    env = default_env_params(num_parallel, num_agents, num_obstacles, sampling_style='policy',
                             device=device)
    env['init'] = dict(TRIANGLE_INIT, init_method='template', num_parallel=num_parallel,
                       num_obs=num_obstacles, device=device,
                       agent_template=agent_template if agent_template is not None
                       else ring_template(num_agents))
    env.update(overrides)
    return env
