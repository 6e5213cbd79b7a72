"""TEST INFRASTRUCTURE -- a mirror of the reference's Python optimizer front-end for this package's
NDArray, kept so the parity tests can drive the store through the reference's CALLBACK route
(``kv._set_updater(get_updater(optimizer))``: the store reduces, then calls the optimizer's own
``update`` per key, which invokes the optimizer operators through MXImperativeInvokeEx) and compare
it with the natively fused route.

It follows python/mxnet/optimizer/optimizer.py of the reference closely on purpose -- :53-524
(Optimizer base: multipliers, update counts), :526-669 (SGD), :775-1030 (LARS), :1547-1629 (Adam),
:1263-1370 (LAMB), :2031-2046 (Test), :2071-2176 (Updater, get_updater) -- and is NOT part of the
product: anand_mxnet_b200/ never imports it. In a real integration the reference's own file plays
this role unchanged.
"""
import math
import os
import pickle
import warnings

import numpy

from anand_mxnet_b200.ndarray import (NDArray, zeros, array, sgd_update, sgd_mom_update, mp_sgd_update,
                      mp_sgd_mom_update, multi_sgd_update, multi_sgd_mom_update, multi_mp_sgd_update,
                      multi_mp_sgd_mom_update, adam_update, cast, multi_sum_sq, multi_lars,
                      preloaded_multi_sgd_update, preloaded_multi_sgd_mom_update,
                      preloaded_multi_mp_sgd_update, preloaded_multi_mp_sgd_mom_update,
                      lamb_update_phase1, lamb_update_phase2, mp_lamb_update_phase1,
                      mp_lamb_update_phase2, contrib as _contrib)

__all__ = ['Optimizer', 'SGD', 'LARS', 'Adam', 'LAMB', 'Test', 'Updater', 'get_updater', 'create',
           'register']


def _flatten_list(nested_list):
    return [item for sublist in nested_list for item in sublist]


class Optimizer(object):
    """The base class inherited by all optimizers (optimizer.py:53-524): hyper-parameters, per-index
    learning-rate / weight-decay multipliers, per-device update counters."""
    opt_registry = {}

    def __init__(self, rescale_grad=1., param_idx2name=None, wd=0., clip_gradient=None,
                 learning_rate=None, lr_scheduler=None, sym=None, begin_num_update=0,
                 multi_precision=False, param_dict=None):
        if lr_scheduler is None:
            learning_rate = 0.01 if learning_rate is None else learning_rate
        elif learning_rate is not None and lr_scheduler.base_lr != learning_rate:
            # an explicit learning rate wins over the scheduler's base (optimizer.py:117-123)
            print(UserWarning("learning rate from ``lr_scheduler`` has been overwritten by "
                              "``learning_rate`` in optimizer."))
            lr_scheduler.base_lr = learning_rate
        assert param_idx2name is None or isinstance(param_idx2name, dict), \
            'param_idx2name should be a dict of param indexes to names.'
        self.lr, self.lr_scheduler, self.wd = learning_rate, lr_scheduler, wd
        self.rescale_grad, self.clip_gradient = rescale_grad, clip_gradient
        self.multi_precision = multi_precision
        self.aggregate_num = 0
        # update counters: one table per device (Trainer shares one optimizer between the updaters
        # of all its devices), `num_update` is the largest count seen anywhere
        self.begin_num_update = self.num_update = begin_num_update
        self._all_index_update_counts = {0: {}}
        self._index_update_count = self._all_index_update_counts[0]
        self.idx2name = dict(param_idx2name or {})
        self.param_dict = param_dict or {}
        self.sym_info = ()
        self.lr_mult, self.wd_mult = {}, {}
        self.set_lr_mult({})
        self.set_wd_mult({})

    @staticmethod
    def register(klass):
        assert isinstance(klass, type)
        key = klass.__name__.lower()
        old = Optimizer.opt_registry.get(key)
        if old is not None:
            warnings.warn('WARNING: New optimizer %s.%s is overriding existing optimizer %s.%s' % (
                klass.__module__, klass.__name__, old.__module__, old.__name__))
        Optimizer.opt_registry[key] = klass
        return klass

    @staticmethod
    def create_optimizer(name, **kwargs):
        klass = Optimizer.opt_registry.get(name.lower())
        if klass is None:
            raise ValueError('Cannot find optimizer %s' % name)
        return klass(**kwargs)

    @property
    def learning_rate(self):
        return self.lr if self.lr_scheduler is None else self.lr_scheduler(self.num_update)

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
        self.lr_mult = dict(args_lr_mult)

    def set_wd_mult(self, args_wd_mult):
        # parameters that are neither *_weight nor *_gamma do not decay unless told otherwise
        self.wd_mult = {n: 0.0 for n in self.idx2name.values()
                        if not n.endswith(('_weight', '_gamma'))}
        self.wd_mult.update(args_wd_mult)

    def _set_current_context(self, device_id):
        self._index_update_count = self._all_index_update_counts.setdefault(device_id, {})

    def _update_count(self, index):
        counts = self._index_update_count
        for idx in (index if isinstance(index, (list, tuple)) else [index]):
            counts[idx] = counts.get(idx, self.begin_num_update) + 1
            if counts[idx] > self.num_update:
                self.num_update = counts[idx]

    def _multiplier(self, table, attr, index):
        """precedence: the Parameter object of the index, then the index, then its name"""
        if index in self.param_dict:
            return getattr(self.param_dict[index], attr)
        if index in table:
            return table[index]
        if index in self.idx2name:
            return table.get(self.idx2name[index], 1.0)
        return None

    def _scaled(self, base, table, attr, indices):
        out = []
        for index in indices:
            m = self._multiplier(table, attr, index)
            out.append(base if m is None else base * m)
        return out

    def _get_lrs(self, indices):
        return self._scaled(self.learning_rate, self.lr_mult, 'lr_mult', indices)

    def _get_lr(self, index):
        return self._get_lrs([index])[0]

    def _get_wds(self, indices):
        return self._scaled(self.wd, self.wd_mult, 'wd_mult', indices)

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


_NO_LARS_SUFFIXES = ('gamma', 'beta', 'bias')


@register
class LARS(Optimizer):
    """Layer-wise adaptive rate scaling on top of SGD-momentum (optimizer.py:798-1055).
    Per layer (except gamma / beta / bias parameters): lr *= eta * |w| / (|g| + wd * |w| + eps)
    when both norms are positive. The dense path never leaves the device: two multi_sum_sq
    launches, one multi_lars over the per-layer lr array, then preloaded_multi_*sgd* updates that
    read lr / wd from device arrays."""

    def __init__(self, momentum=0.0, lazy_update=True, eta=0.001, eps=0, momentum_correction=True,
                 **kwargs):
        super(LARS, self).__init__(**kwargs)
        self.momentum = momentum
        self.momentum_correction = momentum_correction
        self.lazy_update = lazy_update
        self.aggregate_num = int(os.getenv('MXNET_OPTIMIZER_AGGREGATION_SIZE', "4"))
        self.eta = eta
        self.eps = eps
        self.skip = 0
        self.last_lr = None
        self.cur_lr = None

    def _get_lrs(self, indices):
        # also remembers the previous global lr for the momentum correction (optimizer.py:842-871)
        lr = self.learning_rate
        self.last_lr = lr if self.cur_lr is None else self.cur_lr
        self.cur_lr = lr
        return self._scaled(lr, self.lr_mult, 'lr_mult', indices)

    def set_wd_mult(self, args_wd_mult):
        # only *_weight parameters decay (optimizer.py:873-886)
        self.wd_mult = {n: 0.0 for n in self.idx2name.values() if not n.endswith('_weight')}
        if self.sym_info:
            attr, arg_names = self.sym_info
            self.wd_mult.update({name: float(attr[name]['__wd_mult__']) for name in arg_names
                                 if name in attr and '__wd_mult__' in attr[name]})
        self.wd_mult.update(args_wd_mult)

    create_state_multi_precision = SGD.create_state_multi_precision

    def create_state(self, index, weight):
        if self.momentum == 0.0:
            return None
        return zeros(weight.shape, weight.context, dtype=weight.dtype)

    def _name(self, i):
        return self.idx2name[i] if i in self.idx2name else str(i)

    def _l2norm(self, v, rescale=False):
        norm = float(v.astype('float32').norm().asnumpy()[0])
        return norm * self.rescale_grad if rescale else norm

    def _get_lars(self, i, weight, g, lr, wd):
        """per-layer learning rate of the non-aggregated (sparse) route (optimizer.py:919-933)"""
        if self._name(i).endswith(_NO_LARS_SUFFIXES):
            return lr
        w_norm = self._l2norm(weight)
        g_norm = self._l2norm(g, rescale=True)
        if w_norm > 0.0 and g_norm > 0.0:
            return self.eta * w_norm / (g_norm + wd * w_norm + self.eps) * lr
        return lr

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
            kwargs['momentum'] = (self.momentum * (self.cur_lr / self.last_lr)
                                  if (self.momentum_correction and self.last_lr != 0)
                                  else self.momentum)
        if self.clip_gradient:
            kwargs['clip_gradient'] = self.clip_gradient
        if not aggregate:
            lrs = [self._get_lars(i, w, g, lr, wd)
                   for (i, w, g, lr, wd) in zip(indices, weights, grads, lrs, wds)]
            for weight, grad, state, lr, wd in zip(weights, grads, states, lrs, wds):
                if not multi_precision:
                    if state is not None:
                        sgd_mom_update(weight, grad, state, out=weight, lazy_update=self.lazy_update,
                                       lr=lr, wd=wd, **kwargs)
                    else:
                        sgd_update(weight, grad, out=weight, lazy_update=self.lazy_update, lr=lr,
                                   wd=wd, **kwargs)
                elif state[0] is not None:
                    mp_sgd_mom_update(weight, grad, state[0], state[1], out=weight, lr=lr, wd=wd,
                                      **kwargs)
                else:
                    mp_sgd_update(weight, grad, state[1], out=weight, lr=lr, wd=wd, **kwargs)
            return
        # layers that get a LARS coefficient first, the rest (gamma / beta / bias) after them
        n = len(indices)
        skip = [self._name(i).endswith(_NO_LARS_SUFFIXES) for i in indices]
        order = [k for k in range(n) if not skip[k]] + [k for k in range(n) if skip[k]]
        nb_lars = n - sum(skip)
        ctx = weights[0].context
        new_lrs = array([lrs[k] for k in order], ctx=ctx, dtype='float32')
        new_wds = array([wds[k] for k in order], ctx=ctx, dtype='float32')
        ws = [weights[k] for k in order]
        gs = [grads[k] for k in order]
        sts = [states[k] for k in order]
        if nb_lars > 0:
            w_sum_sq = multi_sum_sq(*ws[:nb_lars], num_arrays=nb_lars)
            g_sum_sq = multi_sum_sq(*gs[:nb_lars], num_arrays=nb_lars)
            multi_lars(new_lrs[:nb_lars], w_sum_sq, g_sum_sq, new_wds[:nb_lars], eta=self.eta,
                       eps=self.eps, rescale_grad=self.rescale_grad, out=new_lrs[:nb_lars])
        for sidx in range(0, n, self.aggregate_num):
            eidx = min(sidx + self.aggregate_num, n)
            w, g, st = ws[sidx:eidx], gs[sidx:eidx], sts[sidx:eidx]
            tail = [new_lrs[sidx:eidx], new_wds[sidx:eidx]]
            if not multi_precision:
                if self.momentum > 0:
                    preloaded_multi_sgd_mom_update(*(_flatten_list(zip(w, g, st)) + tail), out=w,
                                                   num_weights=len(w), **kwargs)
                else:
                    preloaded_multi_sgd_update(*(_flatten_list(zip(w, g)) + tail), out=w,
                                               num_weights=len(w), **kwargs)
            elif self.momentum > 0:
                preloaded_multi_mp_sgd_mom_update(*(_flatten_list(zip(w, g, *zip(*st))) + tail),
                                                  out=w, num_weights=len(w), **kwargs)
            else:
                preloaded_multi_mp_sgd_update(*(_flatten_list(zip(w, g, list(zip(*st))[1])) + tail),
                                              out=w, num_weights=len(w), **kwargs)

    update = SGD.update
    update_multi_precision = SGD.update_multi_precision


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
class LAMB(Optimizer):
    """LAMB (optimizer.py:1251-1370): Adam moments, then a per-layer trust ratio |w| / |g'|.
    Lists of tensors take the fused _multi_[mp_]lamb_update operator (<= 45 tensors per call);
    single tensors the two-phase operators with the norms computed in between."""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-6, lower_bound=None,
                 upper_bound=None, bias_correction=True, **kwargs):
        super(LAMB, self).__init__(learning_rate=learning_rate, **kwargs)
        self.beta1 = beta1
        self.beta2 = beta2
        self.epsilon = epsilon
        self.lower_bound = lower_bound
        self.upper_bound = upper_bound
        self.bias_correction = bias_correction
        self.aggregate_num = max(1, min(45, int(os.getenv('MXNET_OPTIMIZER_AGGREGATION_SIZE', "45"))))

    def create_state(self, index, weight):
        return (zeros(weight.shape, weight.context, dtype=weight.dtype),
                zeros(weight.shape, weight.context, dtype=weight.dtype))

    def _bounds(self):
        kw = {}
        if self.lower_bound:
            kw['lower_bound'] = self.lower_bound
        if self.upper_bound:
            kw['upper_bound'] = self.upper_bound
        return kw

    def _update_impl(self, index, weight, grad, state, multi_precision=False):
        kwargs = {'beta1': self.beta1, 'beta2': self.beta2, 'epsilon': self.epsilon,
                  'bias_correction': self.bias_correction, 'rescale_grad': self.rescale_grad}
        if self.clip_gradient:
            kwargs['clip_gradient'] = self.clip_gradient
        if self.aggregate_num <= 1 or not isinstance(index, (tuple, list)):
            if isinstance(index, (tuple, list)):
                assert len(index) == self.aggregate_num
                index, weight, grad, state = index[0], weight[0], grad[0], state[0]
            assert isinstance(weight, NDArray)
            assert isinstance(grad, NDArray)
            self._update_count(index)
            lr = self._get_lr(index)
            wd = self._get_wd(index)
            kwargs['t'] = self._index_update_count[index]
            if multi_precision:
                weight32, (mean, var) = state[0], state[1]
                g = mp_lamb_update_phase1(weight, grad, mean, var, weight32, wd=wd, **kwargs)
                mp_lamb_update_phase2(weight, g, weight32.norm(), g.norm(), weight32, lr=lr,
                                      out=weight, **self._bounds())
            else:
                mean, var = state
                g = lamb_update_phase1(weight, grad, mean, var, wd=wd, **kwargs)
                lamb_update_phase2(weight, g, weight.norm(), g.norm(), lr=lr, out=weight,
                                   **self._bounds())
            return
        kwargs.update(self._bounds())
        step_count, lrs, wds = [], [], []
        for i, w_i, g_i in zip(index, weight, grad):
            assert isinstance(w_i, NDArray)
            assert isinstance(g_i, NDArray)
            self._update_count(i)
            step_count.append(self._index_update_count[i])
            lrs.append(self._get_lr(i))
            wds.append(self._get_wd(i))
        for sidx in range(0, len(weight), self.aggregate_num):
            eidx = min(sidx + self.aggregate_num, len(weight))
            sl = slice(sidx, eidx)
            if not multi_precision:
                mean, var = list(zip(*state[sl]))
                _contrib.multi_lamb_update(weight[sl], grad[sl], mean, var, out=weight[sl],
                                           step_count=step_count[sl], lrs=lrs[sl], wds=wds[sl],
                                           **kwargs)
            else:
                weights32, mean_var = list(zip(*state[sl]))
                mean, var = list(zip(*mean_var))
                _contrib.multi_mp_lamb_update(weight[sl], grad[sl], mean, var, weights32,
                                              out=weight[sl], step_count=step_count[sl],
                                              lrs=lrs[sl], wds=wds[sl], **kwargs)

    def update(self, index, weight, grad, state):
        self._update_impl(index, weight, grad, state, multi_precision=False)

    def update_multi_precision(self, index, weight, grad, state):
        if not isinstance(index, (tuple, list)):
            use_mp = self.multi_precision and weight.dtype == numpy.float16
        else:
            use_mp = self.multi_precision and weight[0].dtype == numpy.float16
        self._update_impl(index, weight, grad, state, multi_precision=use_mp)


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
    """Updater for kvstore (optimizer.py:2071-2161): owns the per-key optimizer states, creates
    them on first sight of a key, and hands keys to the optimizer one at a time or -- when the
    optimizer aggregates -- grouped by dtype in chunks of `aggregate_num`."""

    def __init__(self, optimizer):
        self.optimizer = optimizer
        self.states = {}
        self.states_synced = {}
        self.aggregate_updates = optimizer.aggregate_num > 0

    def _state_for(self, idx, weight):
        if idx not in self.states:
            self.states[idx] = self.optimizer.create_state_multi_precision(idx, weight)
        elif not self.states_synced[idx]:
            # states restored from a checkpoint live wherever they were saved
            self.states[idx] = self.sync_state_context(self.states[idx], weight.context)
        self.states_synced[idx] = True
        return self.states[idx]

    def __call__(self, index, grad, weight):
        single = not isinstance(index, (list, tuple))
        indices = [index] if single else list(index)
        grads = [grad] if single else grad
        weights = [weight] if single else weight
        if weights:
            self.optimizer._set_current_context(weights[0].context.device_id)
        if any(isinstance(i, bytes) for i in indices):
            indices = [i.decode('utf-8') if isinstance(i, bytes) else i for i in indices]
            if isinstance(index, list):
                index[:] = indices           # the reference decodes byte keys in the caller's list
        for idx, w in zip(indices, weights):
            self._state_for(idx, w)
        if not self.aggregate_updates:
            for idx, w, g in zip(indices, weights, grads):
                self.optimizer.update_multi_precision(idx, w, g, self.states[idx])
            return
        by_dtype = {}
        for triple in zip(indices, weights, grads):
            by_dtype.setdefault(str(triple[1].dtype), []).append(triple)
        step = self.optimizer.aggregate_num
        for members in by_dtype.values():
            for at in range(0, len(members), step):
                idxs, ws, gs = (list(x) for x in zip(*members[at:at + step]))
                self.optimizer.update_multi_precision(idxs, ws, gs, [self.states[i] for i in idxs])

    def sync_state_context(self, state, context):
        if isinstance(state, NDArray):
            return state.as_in_context(context)
        if isinstance(state, (tuple, list)):
            return type(state)(self.sync_state_context(i, context) for i in state)
        return state

    def set_states(self, states):
        loaded = pickle.loads(states)
        if isinstance(loaded, tuple) and len(loaded) == 2:
            self.states, self.optimizer = loaded
        else:
            self.states = loaded
        self.states_synced = dict.fromkeys(self.states.keys(), False)

    def get_states(self, dump_optimizer=False):
        return pickle.dumps((self.states, self.optimizer) if dump_optimizer else self.states)


def get_updater(optimizer):
    return Updater(optimizer)
