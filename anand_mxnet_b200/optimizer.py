"""Optimizer front-end for the KVStore path: Optimizer (lr / wd multipliers, update counts), SGD,
Adam, Test and the Updater that a store calls back -- the semantics of
python/mxnet/optimizer/optimizer.py:53-524 (base), :526-669 (SGD), :1547-1629 (Adam),
:2031-2046 (Test), :2071-2176 (Updater) restated for this package.

Two execution routes share this bookkeeping:
  * callback route (any Optimizer subclass): the store reduces, then calls the Updater, whose
    ``update`` runs the reference's optimizer operators through MXImperativeInvokeEx;
  * fused route (SGD / Adam / Test on a KVStore): ``KVStore.set_optimizer`` hands the
    hyper-parameters to the library, which replays exactly this bookkeeping natively and runs the
    step inside the reduce kernel. Results are bit-identical between the routes.
"""
import math
import os
import pickle
import warnings

import numpy

from .ndarray import (NDArray, zeros, sgd_update, sgd_mom_update, mp_sgd_update, mp_sgd_mom_update,
                      multi_sgd_update, multi_sgd_mom_update, multi_mp_sgd_update,
                      multi_mp_sgd_mom_update, adam_update, cast)

__all__ = ['Optimizer', 'SGD', 'Adam', 'Test', 'Updater', 'get_updater', 'create', 'register']


def _flatten_list(nested_list):
    return [item for sublist in nested_list for item in sublist]


class Optimizer(object):
    """The base class inherited by all optimizers (optimizer.py:53-524)."""
    opt_registry = {}

    def __init__(self, rescale_grad=1., param_idx2name=None, wd=0., clip_gradient=None,
                 learning_rate=None, lr_scheduler=None, sym=None, begin_num_update=0,
                 multi_precision=False, param_dict=None):
        self.rescale_grad = rescale_grad
        self.lr_scheduler = lr_scheduler
        if self.lr_scheduler is None and learning_rate is None:
            learning_rate = 0.01
        self.lr = learning_rate
        if self.lr_scheduler is not None and learning_rate is not None:
            if self.lr_scheduler.base_lr != learning_rate:
                print(UserWarning("learning rate from ``lr_scheduler`` has been overwritten by "
                                  "``learning_rate`` in optimizer."))
                self.lr_scheduler.base_lr = learning_rate
        self.wd = wd
        self.lr_mult = {}
        self.wd_mult = {}
        self.begin_num_update = begin_num_update
        self.num_update = begin_num_update
        self._all_index_update_counts = {0: {}}
        self._index_update_count = self._all_index_update_counts[0]
        self.clip_gradient = clip_gradient
        self.multi_precision = multi_precision
        self.aggregate_num = 0
        if param_idx2name is None:
            param_idx2name = {}
        assert isinstance(param_idx2name, dict), \
            'param_idx2name should be a dict of param indexes to names.'
        self.idx2name = param_idx2name.copy()
        self.sym_info = ()
        self.param_dict = param_dict if param_dict else {}
        self.set_lr_mult({})
        self.set_wd_mult({})

    @staticmethod
    def register(klass):
        assert isinstance(klass, type)
        name = klass.__name__.lower()
        if name in Optimizer.opt_registry:
            warnings.warn('WARNING: New optimizer %s.%s is overriding existing optimizer %s.%s' % (
                klass.__module__, klass.__name__, Optimizer.opt_registry[name].__module__,
                Optimizer.opt_registry[name].__name__))
        Optimizer.opt_registry[name] = klass
        return klass

    @staticmethod
    def create_optimizer(name, **kwargs):
        if name.lower() in Optimizer.opt_registry:
            return Optimizer.opt_registry[name.lower()](**kwargs)
        raise ValueError('Cannot find optimizer %s' % name)

    @property
    def learning_rate(self):
        if self.lr_scheduler is not None:
            return self.lr_scheduler(self.num_update)
        return self.lr

    def create_state(self, index, weight):
        """Creates auxiliary state for a given weight."""

    def create_state_multi_precision(self, index, weight):
        if self.multi_precision and weight.dtype == numpy.float16:
            weight_master_copy = weight.astype(numpy.float32)
            return (weight_master_copy,) + (self.create_state(index, weight_master_copy),)
        if weight.dtype == numpy.float16 and not self.multi_precision:
            warnings.warn("Accumulating with float16 in optimizer can lead to poor accuracy or slow "
                          "convergence. Consider using multi_precision=True option of the optimizer")
        return self.create_state(index, weight)

    def update(self, index, weight, grad, state):
        raise NotImplementedError()

    def update_multi_precision(self, index, weight, grad, state):
        if self.multi_precision and weight.dtype == numpy.float16:
            weight_master_copy, original_state = state[0], state[1]
            grad32 = grad.astype(numpy.float32)
            self.update(index, weight_master_copy, grad32, original_state)
            cast(weight_master_copy, dtype=numpy.dtype(weight.dtype).name, out=weight)
        else:
            self.update(index, weight, grad, state)

    def set_learning_rate(self, lr):
        if self.lr_scheduler is not None:
            raise UserWarning("LRScheduler of the optimizer has already been defined. Note that "
                              "set_learning_rate can mutate the value of the learning rate of the "
                              "optimizer only when the LRScheduler of the optimizer is undefined.")
        self.lr = lr

    def set_lr_mult(self, args_lr_mult):
        self.lr_mult = {}
        self.lr_mult.update(args_lr_mult)

    def set_wd_mult(self, args_wd_mult):
        self.wd_mult = {}
        for n in self.idx2name.values():
            if not (n.endswith('_weight') or n.endswith('_gamma')):
                self.wd_mult[n] = 0.0
        self.wd_mult.update(args_wd_mult)

    def _set_current_context(self, device_id):
        if device_id not in self._all_index_update_counts:
            self._all_index_update_counts[device_id] = {}
        self._index_update_count = self._all_index_update_counts[device_id]

    def _update_count(self, index):
        if not isinstance(index, (list, tuple)):
            index = [index]
        for idx in index:
            if idx not in self._index_update_count:
                self._index_update_count[idx] = self.begin_num_update
            self._index_update_count[idx] += 1
            self.num_update = max(self._index_update_count[idx], self.num_update)

    def _get_lrs(self, indices):
        lr = self.lr_scheduler(self.num_update) if self.lr_scheduler is not None else self.lr
        lrs = [lr for _ in indices]
        for i, index in enumerate(indices):
            if index in self.param_dict:
                lrs[i] *= self.param_dict[index].lr_mult
            elif index in self.lr_mult:
                lrs[i] *= self.lr_mult[index]
            elif index in self.idx2name:
                lrs[i] *= self.lr_mult.get(self.idx2name[index], 1.0)
        return lrs

    def _get_lr(self, index):
        return self._get_lrs([index])[0]

    def _get_wds(self, indices):
        wds = [self.wd for _ in indices]
        for i, index in enumerate(indices):
            if index in self.param_dict:
                wds[i] *= self.param_dict[index].wd_mult
            elif index in self.wd_mult:
                wds[i] *= self.wd_mult[index]
            elif index in self.idx2name:
                wds[i] *= self.wd_mult.get(self.idx2name[index], 1.0)
        return wds

    def _get_wd(self, index):
        return self._get_wds([index])[0]

    def __getstate__(self):
        ret = self.__dict__.copy()
        del ret['param_dict']
        return ret

    def __setstate__(self, state):
        self.__dict__ = state
        self.param_dict = {}


register = Optimizer.register
create = Optimizer.create_optimizer


@register
class SGD(Optimizer):
    """SGD with momentum and weight decay (optimizer.py:526-669); dense updates always take the
    multi_*sgd*_update operators, row_sparse ones the (lazy) single-tensor operators."""

    def __init__(self, momentum=0.0, lazy_update=True, **kwargs):
        super(SGD, self).__init__(**kwargs)
        self.momentum = momentum
        self.lazy_update = lazy_update
        self.aggregate_num = int(os.getenv('MXNET_OPTIMIZER_AGGREGATION_SIZE', "4"))

    def create_state_multi_precision(self, index, weight):
        if self.multi_precision and weight.dtype == numpy.float16:
            weight_master_copy = weight.astype(numpy.float32)
            return (self.create_state(index, weight_master_copy), weight_master_copy)
        if weight.dtype == numpy.float16 and not self.multi_precision:
            warnings.warn("Accumulating with float16 in optimizer can lead to poor accuracy or slow "
                          "convergence. Consider using multi_precision=True option of the SGD "
                          "optimizer")
        return self.create_state(index, weight)

    def create_state(self, index, weight):
        momentum = None
        if self.momentum != 0.0:
            momentum = zeros(weight.shape, weight.context, dtype=weight.dtype)
        return momentum

    def _update_impl(self, indices, weights, grads, states, multi_precision=False):
        aggregate = True
        if not isinstance(indices, (tuple, list)):
            indices, weights, grads, states = [indices], [weights], [grads], [states]
        for weight, grad in zip(weights, grads):
            assert isinstance(weight, NDArray)
            assert isinstance(grad, NDArray)
            aggregate = aggregate and weight.stype == 'default' and grad.stype == 'default'
        self._update_count(indices)
        lrs = self._get_lrs(indices)
        wds = self._get_wds(indices)
        kwargs = {'rescale_grad': self.rescale_grad}
        if self.momentum > 0:
            kwargs['momentum'] = self.momentum
        if self.clip_gradient:
            kwargs['clip_gradient'] = self.clip_gradient
        if aggregate:
            if not multi_precision:
                if self.momentum > 0:
                    multi_sgd_mom_update(*_flatten_list(zip(weights, grads, states)), out=weights,
                                         num_weights=len(weights), lrs=tuple(lrs), wds=tuple(wds),
                                         **kwargs)
                else:
                    multi_sgd_update(*_flatten_list(zip(weights, grads)), out=weights,
                                     num_weights=len(weights), lrs=tuple(lrs), wds=tuple(wds),
                                     **kwargs)
            else:
                if self.momentum > 0:
                    multi_mp_sgd_mom_update(*_flatten_list(zip(weights, grads, *zip(*states))),
                                            out=weights, num_weights=len(weights), lrs=tuple(lrs),
                                            wds=tuple(wds), **kwargs)
                else:
                    multi_mp_sgd_update(*_flatten_list(zip(weights, grads, list(zip(*states))[1])),
                                        out=weights, num_weights=len(weights), lrs=tuple(lrs),
                                        wds=tuple(wds), **kwargs)
        else:
            for weight, grad, state, lr, wd in zip(weights, grads, states, lrs, wds):
                if not multi_precision:
                    if state is not None:
                        sgd_mom_update(weight, grad, state, out=weight,
                                       lazy_update=self.lazy_update, lr=lr, wd=wd, **kwargs)
                    else:
                        sgd_update(weight, grad, out=weight, lazy_update=self.lazy_update, lr=lr,
                                   wd=wd, **kwargs)
                else:
                    if state[0] is not None:
                        mp_sgd_mom_update(weight, grad, state[0], state[1], out=weight, lr=lr,
                                          wd=wd, **kwargs)
                    else:
                        mp_sgd_update(weight, grad, state[1], out=weight, lr=lr, wd=wd, **kwargs)

    def update(self, index, weight, grad, state):
        self._update_impl(index, weight, grad, state, multi_precision=False)

    def update_multi_precision(self, index, weight, grad, state):
        if not isinstance(index, (tuple, list)):
            use_mp = self.multi_precision and weight.dtype == numpy.float16
        else:
            use_mp = self.multi_precision and weight[0].dtype == numpy.float16
        self._update_impl(index, weight, grad, state, multi_precision=use_mp)


@register
class Adam(Optimizer):
    """Adam (optimizer.py:1547-1629): bias correction folded into lr in python double."""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, lazy_update=True,
                 **kwargs):
        super(Adam, self).__init__(learning_rate=learning_rate, **kwargs)
        self.beta1 = beta1
        self.beta2 = beta2
        self.epsilon = epsilon
        self.lazy_update = lazy_update

    def create_state(self, index, weight):
        return (zeros(weight.shape, weight.context, dtype=weight.dtype),
                zeros(weight.shape, weight.context, dtype=weight.dtype))

    def update(self, index, weight, grad, state):
        assert isinstance(weight, NDArray)
        assert isinstance(grad, NDArray)
        self._update_count(index)
        lr = self._get_lr(index)
        wd = self._get_wd(index)
        t = self._index_update_count[index]
        coef1 = 1. - self.beta1 ** t
        coef2 = 1. - self.beta2 ** t
        lr *= math.sqrt(coef2) / coef1
        kwargs = {'beta1': self.beta1, 'beta2': self.beta2, 'epsilon': self.epsilon,
                  'rescale_grad': self.rescale_grad}
        if self.clip_gradient:
            kwargs['clip_gradient'] = self.clip_gradient
        mean, var = state
        adam_update(weight, grad, mean, var, out=weight, lazy_update=self.lazy_update, lr=lr, wd=wd,
                    **kwargs)


@register
class Test(Optimizer):
    """The Test optimizer (optimizer.py:2031-2046): w += rescale_grad * g; state mirrors w."""

    def __init__(self, **kwargs):
        super(Test, self).__init__(**kwargs)

    def create_state(self, index, weight):
        return zeros(weight.shape, weight.context)

    def update(self, index, weight, grad, state):
        weight += grad * self.rescale_grad
        state[:] = weight


class Updater(object):
    """Updater for kvstore (optimizer.py:2071-2161)."""

    def __init__(self, optimizer):
        self.optimizer = optimizer
        self.states = {}
        self.states_synced = {}
        self.aggregate_updates = optimizer.aggregate_num > 0

    def __call__(self, index, grad, weight):
        if not isinstance(index, (list, tuple)):
            indices, grads, weights = [index], [grad], [weight]
        else:
            indices, grads, weights = index, grad, weight
        if weights:
            self.optimizer._set_current_context(weights[0].context.device_id)
        for i, idx in enumerate(indices):
            if isinstance(idx, bytes):
                indices[i] = idx.decode('utf-8')
                idx = indices[i]
            if idx not in self.states:
                self.states[idx] = self.optimizer.create_state_multi_precision(idx, weights[i])
                self.states_synced[idx] = True
            elif not self.states_synced[idx]:
                self.states[idx] = self.sync_state_context(self.states[idx], weights[i].context)
                self.states_synced[idx] = True
        if self.aggregate_updates:
            type_map = {}
            for i, w, g in zip(indices, weights, grads):
                type_map.setdefault(str(w.dtype), []).append((i, w, g))
            for idx in type_map:
                current_index = 0
                indices, weights, grads = zip(*type_map[idx])
                while current_index < len(indices):
                    states = []
                    step = min(self.optimizer.aggregate_num, len(indices) - current_index)
                    for j in range(step):
                        states.append(self.states[indices[current_index + j]])
                    n = self.optimizer.aggregate_num
                    self.optimizer.update_multi_precision(
                        list(indices[current_index:current_index + n]),
                        list(weights[current_index:current_index + n]),
                        list(grads[current_index:current_index + n]), states)
                    current_index += n
        else:
            for i, w, g in zip(indices, weights, grads):
                self.optimizer.update_multi_precision(i, w, g, self.states[i])

    def sync_state_context(self, state, context):
        if isinstance(state, NDArray):
            return state.as_in_context(context)
        if isinstance(state, (tuple, list)):
            synced = (self.sync_state_context(i, context) for i in state)
            return tuple(synced) if isinstance(state, tuple) else list(synced)
        return state

    def set_states(self, states):
        states = pickle.loads(states)
        if isinstance(states, tuple) and len(states) == 2:
            self.states, self.optimizer = states
        else:
            self.states = states
        self.states_synced = dict.fromkeys(self.states.keys(), False)

    def get_states(self, dump_optimizer=False):
        return pickle.dumps((self.states, self.optimizer) if dump_optimizer else self.states)


def get_updater(optimizer):
    return Updater(optimizer)
