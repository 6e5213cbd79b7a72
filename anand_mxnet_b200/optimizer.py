"""Optimizer settings for a KVStore whose update step runs INSIDE the reduce kernel.

The reference's ``kv.set_optimizer(optimizer)`` pickles a Python ``Optimizer`` and the store calls
its ``update`` back once per key (python/mxnet/kvstore/kvstore.py:543-590,
python/mxnet/optimizer/optimizer.py:2071-2176). Here the step for SGD / Adam / Test is part of the
fused sm_100a kernel, so what the store needs from Python is only a RECORD of hyper-parameters:

    kv.set_optimizer(mx.optimizer.SGD(learning_rate=0.1, momentum=0.9, wd=1e-4, rescale_grad=1/256))
    kv.set_optimizer(mx.optimizer.create('adam', learning_rate=1e-3))

``FusedOptimizer`` is that record: the scalar hyper-parameters, the per-parameter learning-rate /
weight-decay multipliers with the reference's lookup order (Parameter object, then index, then
name; names that are neither ``*_weight`` nor ``*_gamma`` do not decay -- optimizer.py:432-509), an
optional ``lr_scheduler``. It holds no update rule. The library replays the reference's
bookkeeping (update counts, Adam bias correction, python-repr -> dmlc::stof scalar plumbing)
natively -- csrc/kvstore_core.cc -- so results equal the reference's callback route bit for bit.

Any OTHER optimizer is an object with ``create_state_multi_precision`` / ``update_multi_precision``
(the reference's Optimizer protocol): ``get_updater`` wraps it into the callback the store invokes
per key after the reduce (MXKVStoreSetUpdaterEx). The reference's own Optimizer classes satisfy the
protocol; tests/compat/mxnet_optimizer.py carries a mirror of them for this package's NDArray.
"""
import pickle

__all__ = ['FusedOptimizer', 'SGD', 'Adam', 'Test', 'create', 'get_updater', 'fused_kind']

_DEFAULTS = {
    'sgd': dict(learning_rate=0.01, momentum=0.0, lazy_update=True),
    'adam': dict(learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, lazy_update=True),
    'test': dict(learning_rate=0.01),
}
_COMMON = dict(rescale_grad=1.0, wd=0.0, clip_gradient=None, lr_scheduler=None, begin_num_update=0,
               multi_precision=False, param_idx2name=None, param_dict=None, sym=None)


class FusedOptimizer(object):
    """Hyper-parameters of an optimizer the library runs natively (``kind``: sgd / adam / test)."""

    def __init__(self, kind, **kwargs):
        kind = kind.lower()
        if kind not in _DEFAULTS:
            raise ValueError("no fused kernel for optimizer '%s' (sgd, adam, test); pass an object with "
                             "an update rule to kv.set_optimizer for the callback route" % kind)
        self.kind = kind
        known = dict(_COMMON)
        known.update(_DEFAULTS[kind])
        unknown = set(kwargs) - set(known)
        if unknown:
            raise TypeError("%s: unexpected arguments %s" % (kind, sorted(unknown)))
        known.update(kwargs)
        sched = known['lr_scheduler']
        lr_given = 'learning_rate' in kwargs and kwargs['learning_rate'] is not None
        if sched is not None:
            if lr_given and sched.base_lr != kwargs['learning_rate']:
                sched.base_lr = kwargs['learning_rate']   # an explicit rate wins (optimizer.py:117-123)
            if not lr_given:
                known['learning_rate'] = None
        self.lr = known.pop('learning_rate')
        self.lr_scheduler = known.pop('lr_scheduler')
        self.idx2name = dict(known.pop('param_idx2name') or {})
        self.param_dict = known.pop('param_dict') or {}
        known.pop('sym')
        for k, v in known.items():           # wd, rescale_grad, clip_gradient, momentum, beta1 ...
            setattr(self, k, v)
        self.num_update = self.begin_num_update
        self.lr_mult = {}
        self.wd_mult = {}
        self.set_wd_mult({})

    # ---- the reference's setters (optimizer.py:338-410)
    @property
    def learning_rate(self):
        return self.lr if self.lr_scheduler is None else self.lr_scheduler(self.num_update)

    def set_learning_rate(self, lr):
        if self.lr_scheduler is not None:
            raise UserWarning("LRScheduler of the optimizer has already been defined. Note that "
                              "set_learning_rate can mutate the value of the learning rate of the "
                              "optimizer only when the LRScheduler of the optimizer is undefined.")
        self.lr = lr

    def set_lr_mult(self, args_lr_mult):
        self.lr_mult = dict(args_lr_mult)

    def set_wd_mult(self, args_wd_mult):
        self.wd_mult = {name: 0.0 for name in self.idx2name.values()
                        if not name.endswith(('_weight', '_gamma'))}
        self.wd_mult.update(args_wd_mult)

    def multipliers(self, index):
        """(lr multiplier, wd multiplier) of one parameter: its Parameter object first, then a table
        entry for the index, then one for its name, else 1."""
        if index in self.param_dict:
            p = self.param_dict[index]
            return p.lr_mult, p.wd_mult
        out = []
        for table in (self.lr_mult, self.wd_mult):
            if index in table:
                out.append(table[index])
            else:
                out.append(table.get(self.idx2name.get(index), 1.0))
        return tuple(out)

    def op_params(self):
        """what B200KVStoreSetOptimizer receives: the reference's operator-parameter names"""
        p = {'learning_rate': self.learning_rate, 'wd': self.wd, 'rescale_grad': self.rescale_grad,
             'clip_gradient': self.clip_gradient if self.clip_gradient else 0.0,
             'multi_precision': bool(self.multi_precision), 'begin_num_update': self.begin_num_update}
        if self.kind == 'sgd':
            p.update(momentum=self.momentum, lazy_update=bool(self.lazy_update))
        elif self.kind == 'adam':
            p.update(beta1=self.beta1, beta2=self.beta2, epsilon=self.epsilon,
                     lazy_update=bool(self.lazy_update))
        return p

    def __getstate__(self):
        state = self.__dict__.copy()
        state['param_dict'] = {}             # Parameter objects do not travel (optimizer.py:512-519)
        return state


def SGD(**kwargs):
    return FusedOptimizer('sgd', **kwargs)


def Adam(**kwargs):
    return FusedOptimizer('adam', **kwargs)


def Test(**kwargs):
    return FusedOptimizer('test', **kwargs)


def create(name, **kwargs):
    return FusedOptimizer(name, **kwargs)


def fused_kind(optimizer):
    """'sgd' / 'adam' / 'test' when the library can run this optimizer inside the reduce kernel:
    a FusedOptimizer record, or an object of the reference's SGD / Adam / Test classes (matched by
    class name and the attributes the native replay needs)."""
    if isinstance(optimizer, FusedOptimizer):
        return optimizer.kind
    kind = {'SGD': 'sgd', 'Adam': 'adam', 'Test': 'test'}.get(type(optimizer).__name__)
    if kind is None:
        return None
    needed = {'sgd': ('momentum', 'lazy_update'), 'adam': ('beta1', 'beta2', 'epsilon'), 'test': ()}[kind]
    base = ('lr', 'wd', 'rescale_grad', 'clip_gradient', 'multi_precision', 'begin_num_update',
            'lr_mult', 'wd_mult', 'idx2name', 'param_dict')
    return kind if all(hasattr(optimizer, a) for a in base + needed) else None


def record_of(optimizer):
    """a FusedOptimizer view of an optimizer object that passed ``fused_kind``"""
    if isinstance(optimizer, FusedOptimizer):
        return optimizer
    kind = fused_kind(optimizer)
    rec = FusedOptimizer.__new__(FusedOptimizer)
    rec.__dict__.update({k: v for k, v in optimizer.__dict__.items() if not k.startswith('_')})
    rec.kind = kind
    rec.lr_scheduler = getattr(optimizer, 'lr_scheduler', None)
    if not hasattr(rec, 'num_update'):
        rec.num_update = rec.begin_num_update
    if kind == 'sgd' and not hasattr(rec, 'lazy_update'):
        rec.lazy_update = True
    rec._source = optimizer
    return rec


class _CallbackUpdater(object):
    """The per-key callback a store invokes after the reduce: state on first use, then the
    optimizer's own update (the role of optimizer.py:2071-2128, without its aggregation)."""

    def __init__(self, optimizer):
        for need in ('create_state_multi_precision', 'update_multi_precision'):
            if not callable(getattr(optimizer, need, None)):
                raise TypeError("kv.set_optimizer: %s has no %s(); the callback route needs an object "
                                "with the reference's Optimizer protocol" % (type(optimizer).__name__, need))
        self.optimizer = optimizer
        self.states = {}

    def __call__(self, index, grad, weight):
        if index not in self.states:
            self.states[index] = self.optimizer.create_state_multi_precision(index, weight)
        self.optimizer.update_multi_precision(index, weight, grad, self.states[index])

    def get_states(self, dump_optimizer=False):
        return pickle.dumps((self.states, self.optimizer) if dump_optimizer else self.states)

    def set_states(self, states):
        loaded = pickle.loads(states)
        if isinstance(loaded, tuple) and len(loaded) == 2:
            self.states, self.optimizer = loaded
        else:
            self.states = loaded


def get_updater(optimizer):
    return _CallbackUpdater(optimizer)
