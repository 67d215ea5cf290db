"""TEST INFRASTRUCTURE - not product code.

Minimal stand-ins for rclpy / tf2_ros / ROS message packages so that the
reference's hot-path modules import UNMODIFIED from /root/reference in this
container.  Used only by the golden-vector generator
(`oracle/make_golden.py`); nothing on the GPU box imports this, because
/root/reference does not exist there.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may
import anything under `oracle/`.
"""
import sys
import time
import types


class _Bag:
    """Permissive attribute bag: auto-creates nested attributes, takes kwargs."""

    def __init__(self, *a, **kw):
        for k, v in kw.items():
            object.__setattr__(self, k, v)

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        v = _Bag()
        object.__setattr__(self, name, v)
        return v


def _msg_class(name, **consts):
    cls = type(name, (_Bag,), dict(consts))
    return cls


class _Logger:
    def info(self, *a, **k):
        pass
    warn = warning = error = debug = info


class _Publisher:
    def __init__(self):
        self.sent = []

    def publish(self, msg):
        self.sent.append(msg)


class _Clock:
    class _Now:
        def to_msg(self):
            t = time.time()
            return _Bag(sec=int(t), nanosec=int((t - int(t)) * 1e9))

    def now(self):
        return self._Now()


class Node:
    def __init__(self, name='node', *a, **k):
        self._name = name
        self._pubs = {}

    def create_subscription(self, *a, **k):
        return None

    def create_timer(self, *a, **k):
        return None

    def create_publisher(self, typ, topic, *a, **k):
        p = _Publisher()
        self._pubs[topic] = p
        return p

    def get_logger(self):
        return _Logger()

    def get_clock(self):
        return _Clock()

    def destroy_node(self):
        pass


class TfBuffer:
    """tf2_ros.Buffer stand-in; set `.current` to (tx,ty,tz,qx,qy,qz,qw)."""

    def __init__(self, *a, **k):
        self.current = None

    def lookup_transform(self, *a, **k):
        if self.current is None:
            raise RuntimeError('no tf')
        tx, ty, tz, qx, qy, qz, qw = self.current
        return _Bag(transform=_Bag(
            translation=_Bag(x=tx, y=ty, z=tz),
            rotation=_Bag(x=qx, y=qy, z=qz, w=qw)))


def install():
    """Register the stub modules in sys.modules (idempotent)."""
    if 'rclpy' in sys.modules and getattr(sys.modules['rclpy'], '_is_stub', False):
        return

    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    rclpy = mod('rclpy', _is_stub=True, init=lambda *a, **k: None,
                shutdown=lambda *a, **k: None, spin=lambda *a, **k: None,
                ok=lambda: True)
    rclpy.node = mod('rclpy.node', Node=Node)
    rclpy.time = mod('rclpy.time', Time=_msg_class('Time'))
    rclpy.executors = mod('rclpy.executors',
                          ExternalShutdownException=type('ExternalShutdownException', (Exception,), {}))
    rclpy.qos = mod('rclpy.qos', QoSProfile=_Bag, ReliabilityPolicy=_Bag(), DurabilityPolicy=_Bag(),
                    HistoryPolicy=_Bag(), qos_profile_sensor_data=_Bag())
    callable_cls = lambda n: type(n, (), {'__init__': lambda self, *a, **k: None,
                                          'sendTransform': lambda self, *a, **k: None})
    mod('tf2_ros', Buffer=TfBuffer, TransformListener=callable_cls('TransformListener'),
        TransformBroadcaster=callable_cls('TransformBroadcaster'),
        StaticTransformBroadcaster=callable_cls('StaticTransformBroadcaster'))
    mod('sensor_msgs')
    mod('sensor_msgs.msg', Image=_msg_class('Image'), CameraInfo=_msg_class('CameraInfo'),
        PointCloud2=_msg_class('PointCloud2'),
        PointField=_msg_class('PointField', FLOAT32=7), Imu=_msg_class('Imu'))
    mod('geometry_msgs')
    mod('geometry_msgs.msg',
        PoseWithCovarianceStamped=_msg_class('PoseWithCovarianceStamped'),
        TransformStamped=_msg_class('TransformStamped'), Quaternion=_msg_class('Quaternion'),
        Twist=_msg_class('Twist'), PoseStamped=_msg_class('PoseStamped'))
    mod('nav_msgs')
    mod('nav_msgs.msg', Odometry=_msg_class('Odometry'), Path=_msg_class('Path'))
    mod('builtin_interfaces')
    mod('builtin_interfaces.msg', Time=_msg_class('Time'))
    mod('std_msgs')
    mod('std_msgs.msg', Header=_msg_class('Header'), String=_msg_class('String'))


REFERENCE_COMMON = '/root/reference/simulation/isaac/scripts/common'
REFERENCE_SELFTEST_DIR = '/root/reference/simulation/isaac/routes/03_south/teach/scripts'


def import_reference():
    """Import the four hot-path modules unmodified. Returns a dict of modules."""
    import importlib
    import os
    if not os.path.isdir(REFERENCE_COMMON):
        raise RuntimeError('/root/reference is not present (expected on the GPU box)')
    install()
    saved_argv = sys.argv
    sys.argv = ['x']
    if REFERENCE_COMMON not in sys.path:
        sys.path.insert(0, REFERENCE_COMMON)
    try:
        mods = {n: importlib.import_module(n) for n in (
            'visual_landmark_matcher', 'teach_run_depth_mapper', 'tf_wall_clock_relay')}
        spec = importlib.util.spec_from_file_location(
            'checkpoint_a_selftest', os.path.join(REFERENCE_SELFTEST_DIR, 'checkpoint_a_selftest.py'))
        st = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(st)
        mods['checkpoint_a_selftest'] = st
    finally:
        sys.argv = saved_argv
    return mods
