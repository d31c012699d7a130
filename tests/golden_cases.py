"""Shared drivers for the golden-fixture tests (tests/golden/*.npz were produced by the reference's own source,
see tests/golden/make_golden.py).  Used by the CPU tier (CUDA binding replaced by tests/raw_standin.py) and by the
GPU tier (real kernels)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False))


def kernel_class(name):
    from mxfusion_b200.components.distributions.gp.kernels import RBF, Matern12, Matern32, Matern52
    return {'rbf': RBF, 'matern12': Matern12, 'matern32': Matern32, 'matern52': Matern52}[str(name)]


def param_grad(infr, var):
    return infr.params.param_dict[var.uuid].tensor.grad.detach().cpu().numpy().copy()


def run_svgp_case(mf, g, i, device):
    """Rebuilds case i of svgp_fixture.npz through the public API; returns (loss, grads wrt stored parameters)."""
    from mxfusion_b200.components.variables import PositiveTransformation
    from mxfusion_b200.modules.gp_modules import SVGPRegression
    from mxfusion_b200.inference import GradBasedInference, MAP, MinibatchInferenceLoop, BatchInferenceLoop
    c = lambda k: g['case%d_%s' % (i, k)]
    X, Y, Z = c('X'), c('Y'), c('Z')
    N, Din = X.shape
    M, P = Z.shape[0], Y.shape[1]
    m = mf.Model()
    m.N = mf.Variable()
    m.X = mf.Variable(shape=(m.N, Din))
    m.Z = mf.Variable(shape=(M, Din), initial_value=Z)
    m.noise_var = mf.Variable(transformation=PositiveTransformation(), initial_value=c('noise_var'))
    kernel = kernel_class(c('kernel'))(input_dim=Din, ARD=True, variance=c('variance'), lengthscale=c('lengthscale'))
    m.Y = SVGPRegression.define_variable(X=m.X, kernel=kernel, noise_var=m.noise_var, inducing_inputs=m.Z,
                                         shape=(m.N, P))
    gp = m.Y.factor
    gp.svgp_log_pdf.jitter = float(c('jitter'))
    sc = float(c('rv_scaling'))
    loop = MinibatchInferenceLoop(batch_size=N, rv_scaling={m.Y: sc}) if sc != 1.0 else BatchInferenceLoop()
    infr = GradBasedInference(MAP(model=m, observed=[m.X, m.Y]), grad_loop=loop, context=device)
    infr.initialize(X=X.shape, Y=Y.shape)
    post = gp._extra_graphs[0]
    infr.params[post.qU_mean] = c('qU_mean')
    infr.params[post.qU_cov_W] = c('qU_cov_W')
    infr.params[post.qU_cov_diag] = c('qU_cov_diag')
    infr.params.gflat.zero_()
    loss, loss_g = infr.create_executor()(None, torch.tensor(X, device=device), torch.tensor(Y, device=device))
    loss_g.backward()
    grads = dict(Z=param_grad(infr, m.Z), noise_var=param_grad(infr, m.noise_var),
                 qU_mean=param_grad(infr, post.qU_mean), qU_cov_W=param_grad(infr, post.qU_cov_W),
                 qU_cov_diag=param_grad(infr, post.qU_cov_diag), lengthscale=param_grad(infr, kernel.lengthscale),
                 variance=param_grad(infr, kernel.variance))
    return float(loss), grads


def run_gp_case(mf, g, i, device):
    from mxfusion_b200.components.variables import PositiveTransformation
    from mxfusion_b200.modules.gp_modules import GPRegression
    from mxfusion_b200.inference import GradBasedInference, MAP
    c = lambda k: g['case%d_%s' % (i, k)]
    X, Y = c('X'), c('Y')
    N, Din = X.shape
    P = Y.shape[1]
    m = mf.Model()
    m.N = mf.Variable()
    m.X = mf.Variable(shape=(m.N, Din))
    m.noise_var = mf.Variable(transformation=PositiveTransformation(), initial_value=c('noise_var'))
    kernel = kernel_class(c('kernel'))(input_dim=Din, ARD=True, variance=c('variance'), lengthscale=c('lengthscale'))
    m.Y = GPRegression.define_variable(X=m.X, kernel=kernel, noise_var=m.noise_var, shape=(m.N, P))
    infr = GradBasedInference(MAP(model=m, observed=[m.X, m.Y]), context=device)
    infr.initialize(X=X.shape, Y=Y.shape)
    infr.params.gflat.zero_()
    loss, loss_g = infr.create_executor()(None, torch.tensor(X, device=device), torch.tensor(Y, device=device))
    loss_g.backward()
    post = m.Y.factor._extra_graphs[0]
    grads = dict(noise_var=param_grad(infr, m.noise_var), lengthscale=param_grad(infr, kernel.lengthscale),
                 variance=param_grad(infr, kernel.variance))
    return float(loss), grads, infr.params[post.L].cpu().numpy(), infr.params[post.LinvY].cpu().numpy()


def run_svgp_minibatch(mf, g, device, **loop_kw):
    from mxfusion_b200.components.variables import PositiveTransformation
    from mxfusion_b200.components.distributions.gp.kernels import RBF
    from mxfusion_b200.modules.gp_modules import SVGPRegression
    from mxfusion_b200.inference import GradBasedInference, MAP, MinibatchInferenceLoop
    N, B, M, epochs = int(g['N']), int(g['B']), int(g['M']), int(g['epochs'])
    X, Y = g['X'], g['Y']
    m = mf.Model()
    m.N = mf.Variable()
    m.X = mf.Variable(shape=(m.N, 1))
    m.noise_var = mf.Variable(shape=(1,), transformation=PositiveTransformation(), initial_value=0.01)
    m.kernel = RBF(input_dim=1, variance=1, lengthscale=1)
    m.Y = SVGPRegression.define_variable(X=m.X, kernel=m.kernel, noise_var=m.noise_var, shape=(m.N, 1),
                                         num_inducing=M)
    m.Y.factor.svgp_log_pdf.jitter = 1e-6
    loop = MinibatchInferenceLoop(batch_size=B, rv_scaling={m.Y: N / B}, **loop_kw)
    infr = GradBasedInference(inference_algorithm=MAP(model=m, observed=[m.X, m.Y]), grad_loop=loop, context=device)
    infr.initialize(X=(N, 1), Y=(N, 1))
    post = m.Y.factor._extra_graphs[0]
    infr.params[m.Y.factor.inducing_inputs] = g['Z0']
    infr.params[post.qU_mean] = np.zeros((M, 1))
    infr.params[post.qU_cov_W] = np.eye(M) * 0.1
    infr.params[post.qU_cov_diag] = np.ones(M) * 0.5
    np.random.seed(int(g['shuffle_seed']))
    infr.run(X=X, Y=Y, max_iter=epochs, learning_rate=float(g['lr']))
    return dict(final_Z=infr.params[m.Y.factor.inducing_inputs], final_qU_mean=infr.params[post.qU_mean],
                final_qU_cov_diag=infr.params[post.qU_cov_diag], final_lengthscale=infr.params[m.kernel.lengthscale],
                final_variance=infr.params[m.kernel.variance], final_noise_var=infr.params[m.noise_var])


def run_svi_toy(mf, g, device):
    from mxfusion_b200.components.variables import PositiveTransformation
    from mxfusion_b200.components.distributions import Normal, MockMXNetRandomGenerator
    from mxfusion_b200.inference import GradBasedInference, StochasticVariationalInference, create_Gaussian_meanfield
    y = g['y']
    N, S = y.shape[0], int(g['S'])
    t = lambda a: torch.tensor(np.atleast_1d(a), dtype=torch.float64)
    m = mf.Model()
    m.mu = Normal.define_variable(mean=t(g['prior_mean']), variance=t(g['prior_var']), shape=(1,))
    m.s2 = mf.Variable(shape=(1,), transformation=PositiveTransformation(), initial_value=t(g['s2']))
    m.y = Normal.define_variable(mean=m.mu, variance=m.s2, shape=(N, 1))
    q = create_Gaussian_meanfield(model=m, observed=[m.y])
    q.mu.factor._rand_gen = MockMXNetRandomGenerator(torch.tensor(g['eps'].flatten(), device=device))
    alg = StochasticVariationalInference(num_samples=S, model=m, posterior=q, observed=[m.y])
    infr = GradBasedInference(inference_algorithm=alg, context=device)
    infr.initialize(y=y.shape)
    infr.params[q.mu.factor.mean] = t(g['q_mean'])
    infr.params[q.mu.factor.variance] = t(g['q_var'])
    infr.params.gflat.zero_()
    loss, loss_g = infr.create_executor()(None, torch.tensor(y, device=device))
    loss_g.backward()
    grads = dict(q_mean=param_grad(infr, q.mu.factor.mean), q_var=param_grad(infr, q.mu.factor.variance),
                 s2=param_grad(infr, m.s2))
    return float(loss), grads


def run_predict(mf, g, module, device):
    """Rebuilds the prediction scenario of predict.npz; returns {(noise_free, diag): (mean, var)}."""
    from mxfusion_b200.components.variables import PositiveTransformation
    from mxfusion_b200.components.distributions.gp.kernels import RBF
    from mxfusion_b200.modules.gp_modules import SVGPRegression, GPRegression
    from mxfusion_b200.inference import Inference, MAP, TransferInference, ModulePredictionAlgorithm
    X, Y, Z, Xt = g['X'], g['Y'], g['Z'], g['Xt']
    N, Din = X.shape
    M, P = Z.shape[0], Y.shape[1]
    m = mf.Model()
    m.N = mf.Variable()
    m.X = mf.Variable(shape=(m.N, Din))
    m.noise_var = mf.Variable(transformation=PositiveTransformation(), initial_value=g['noise_var'])
    kernel = RBF(input_dim=Din, ARD=True, variance=g['variance'], lengthscale=g['lengthscale'])
    if module == 'svgp':
        m.Z = mf.Variable(shape=(M, Din), initial_value=Z)
        m.Y = SVGPRegression.define_variable(X=m.X, kernel=kernel, noise_var=m.noise_var, inducing_inputs=m.Z,
                                             shape=(m.N, P))
        m.Y.factor.svgp_log_pdf.jitter = 1e-8
    else:
        m.Y = GPRegression.define_variable(X=m.X, kernel=kernel, noise_var=m.noise_var, shape=(m.N, P))
    gp = m.Y.factor
    infr = Inference(MAP(model=m, observed=[m.X, m.Y]), context=device)
    infr.initialize(X=X.shape, Y=Y.shape)
    if module == 'svgp':
        post = gp._extra_graphs[0]
        infr.params[post.qU_mean] = g['qU_mean']
        infr.params[post.qU_cov_W] = g['qU_cov_W']
        infr.params[post.qU_cov_diag] = g['qU_cov_diag']
    infr.run(X=X, Y=Y)
    alg = gp.svgp_predict if module == 'svgp' else gp.gp_predict
    out = {}
    for noise_free in (True, False):
        for diag in (True, False):
            alg.noise_free, alg.diagonal_variance = noise_free, diag
            infr2 = TransferInference(ModulePredictionAlgorithm(m, observed=[m.X], target_variables=[m.Y]),
                                      infr_params=infr.params, context=device)
            with torch.no_grad():
                res = infr2.run(X=Xt)[0]
            out[(noise_free, diag)] = (res[0].cpu().numpy(), res[1].cpu().numpy())
    return out


def run_sparsegp_case(mf, g, i, device, chunk_rows=None):
    """Rebuilds case i of sparsegp_fixture.npz through the public API; returns (loss, grads wrt the stored
    parameters, cached (wv, L, LA), {(noise_free, diag): (mean, var)})."""
    from mxfusion_b200.components.variables import PositiveTransformation
    from mxfusion_b200.modules.gp_modules import SparseGPRegression
    from mxfusion_b200.inference import GradBasedInference, MAP, TransferInference, ModulePredictionAlgorithm
    c = lambda k: g['case%d_%s' % (i, k)]
    X, Y, Z, Xt = c('X'), c('Y'), c('Z'), c('Xt')
    N, Din = X.shape
    M, P = Z.shape[0], Y.shape[1]
    ard = c('lengthscale').shape[0] == Din and Din > 1
    m = mf.Model()
    m.N = mf.Variable()
    m.X = mf.Variable(shape=(m.N, Din))
    m.Z = mf.Variable(shape=(M, Din), initial_value=Z)
    m.noise_var = mf.Variable(transformation=PositiveTransformation(), initial_value=c('noise_var'))
    kernel = kernel_class(c('kernel'))(input_dim=Din, ARD=ard, variance=c('variance'), lengthscale=c('lengthscale'))
    m.Y = SparseGPRegression.define_variable(X=m.X, kernel=kernel, noise_var=m.noise_var, inducing_inputs=m.Z,
                                             shape=(m.N, P))
    gp = m.Y.factor
    gp.sgp_log_pdf.jitter = float(c('jitter'))
    gp.sgp_log_pdf.chunk_rows = chunk_rows
    infr = GradBasedInference(MAP(model=m, observed=[m.X, m.Y]), context=device)
    infr.initialize(X=X.shape, Y=Y.shape)
    infr.params.gflat.zero_()
    loss, loss_g = infr.create_executor()(None, torch.tensor(X, device=device), torch.tensor(Y, device=device))
    loss_g.backward()
    grads = dict(Z=param_grad(infr, m.Z), noise_var=param_grad(infr, m.noise_var),
                 lengthscale=param_grad(infr, kernel.lengthscale), variance=param_grad(infr, kernel.variance))
    post = gp._extra_graphs[0]
    cache = tuple(infr.params[v].cpu().numpy() for v in (post.wv, post.L, post.LA))
    pred = {}
    for noise_free in (True, False):
        for diag in (True, False):
            gp.sgp_predict.noise_free, gp.sgp_predict.diagonal_variance = noise_free, diag
            infr2 = TransferInference(ModulePredictionAlgorithm(m, observed=[m.X], target_variables=[m.Y]),
                                      infr_params=infr.params, context=device)
            with torch.no_grad():
                res = infr2.run(X=Xt)[0]
            pred[(noise_free, diag)] = (res[0].cpu().numpy(), res[1].cpu().numpy())
    return float(loss), grads, cache, pred


# ------------------------------------------------------------------------------------------------ kernel algebra (8f-4)
def combo_kernel(spec, Din):
    """Same constructions as tests/golden/make_golden.py::combo_kernel, with this package's classes."""
    from mxfusion_b200.components.distributions.gp.kernels import RBF, Matern32, Matern52, Linear, Bias, White
    if spec == 'linear':
        return Linear(input_dim=Din, ARD=False)
    if spec == 'linear_ard':
        return Linear(input_dim=Din, ARD=True)
    if spec == 'bias':
        return Bias(input_dim=Din)
    if spec == 'white':
        return White(input_dim=Din)
    if spec == 'rbf+linear_ard':
        return RBF(input_dim=Din, ARD=True) + Linear(input_dim=Din, ARD=True)
    if spec == 'rbf*matern32':
        return RBF(input_dim=Din, ARD=False) * Matern32(input_dim=Din, ARD=True)
    if spec == 'rbf+rbf+bias':
        return RBF(input_dim=Din) + RBF(input_dim=Din, ARD=True) + Bias(input_dim=Din)
    if spec == '(matern52+white)*linear':
        return (Matern52(input_dim=Din) + White(input_dim=Din)) * Linear(input_dim=Din)
    raise ValueError(spec)


def run_combo_kernels(g, device):
    """Yields (tag, got, want) for K(X), K(X, X2), Kdiag of every kernel spec / sample count in combo_kernels.npz."""
    from mxfusion_b200 import F
    for i, spec in enumerate([str(s) for s in g['specs']]):
        for S in (1, 2):
            tag = 'k%d_S%d' % (i, S)
            X, X2 = g[tag + '_X'], g[tag + '_X2']
            k = combo_kernel(spec, X.shape[-1])
            names = [str(n) for n in g[tag + '_names']]
            assert sorted(k.parameters.keys()) == names, (spec, sorted(k.parameters.keys()), names)
            T = lambda a: torch.tensor(a, device=device)
            params = {n: T(g[tag + '_p_' + n]) for n in names}
            yield tag + ' ' + spec + ' K', k.K(F, T(X), **params).cpu().numpy(), g[tag + '_K']
            yield tag + ' ' + spec + ' K2', k.K(F, T(X), T(X2), **params).cpu().numpy(), g[tag + '_K2']
            yield tag + ' ' + spec + ' Kdiag', k.Kdiag(F, T(X), **params).cpu().numpy(), g[tag + '_Kdiag']


def run_combo_module_case(mf, g, i, device):
    """Rebuilds case i of combo_modules.npz (a GP module over a combination kernel); returns (loss, grads)."""
    from mxfusion_b200.components.variables import PositiveTransformation
    from mxfusion_b200.modules.gp_modules import GPRegression, SVGPRegression, SparseGPRegression
    from mxfusion_b200.inference import GradBasedInference, MAP
    c = lambda k: g['case%d_%s' % (i, k)]
    module, spec = str(c('module')), str(c('spec'))
    X, Y, Z = c('X'), c('Y'), c('Z')
    N, Din = X.shape
    M, P = Z.shape[0], Y.shape[1]
    m = mf.Model()
    m.N = mf.Variable()
    m.X = mf.Variable(shape=(m.N, Din))
    m.noise_var = mf.Variable(transformation=PositiveTransformation(), initial_value=c('noise_var'))
    kernel = combo_kernel(spec, Din)
    if module == 'gp':
        m.Y = GPRegression.define_variable(X=m.X, kernel=kernel, noise_var=m.noise_var, shape=(m.N, P))
    else:
        m.Z = mf.Variable(shape=(M, Din), initial_value=Z)
        cls = SVGPRegression if module == 'svgp' else SparseGPRegression
        m.Y = cls.define_variable(X=m.X, kernel=kernel, noise_var=m.noise_var, inducing_inputs=m.Z, shape=(m.N, P))
        (m.Y.factor.svgp_log_pdf if module == 'svgp' else m.Y.factor.sgp_log_pdf).jitter = 1e-6
    gp = m.Y.factor
    infr = GradBasedInference(MAP(model=m, observed=[m.X, m.Y]), context=device)
    infr.initialize(X=X.shape, Y=Y.shape)
    names = [str(n) for n in c('names')]
    for n in names:
        infr.params[kernel.parameters[n]] = c('p_' + n)
    gvars = {'noise_var': m.noise_var}
    if module != 'gp':
        gvars['Z'] = m.Z
    if module == 'svgp':
        post = gp._extra_graphs[0]
        for nm in ('qU_mean', 'qU_cov_W', 'qU_cov_diag'):
            infr.params[getattr(post, nm)] = c(nm)
            gvars[nm] = getattr(post, nm)
    for n in names:
        gvars['p_' + n] = kernel.parameters[n]
    infr.params.gflat.zero_()
    loss, loss_g = infr.create_executor()(None, torch.tensor(X, device=device), torch.tensor(Y, device=device))
    loss_g.backward()
    return float(loss), {k: param_grad(infr, v) for k, v in gvars.items()}


# ------------------------------------------------------------------------------------------------ GP distributions (8f-3)
def run_gp_distributions(g, device):
    """Yields (tag, got, want) for log-pdf and injected-noise draws of GaussianProcess / ConditionalGaussianProcess."""
    from mxfusion_b200 import F, Variable
    from mxfusion_b200.components.distributions import GaussianProcess, ConditionalGaussianProcess, \
        MockMXNetRandomGenerator
    T = lambda a: torch.tensor(a, device=device)
    for i in range(int(g['n_cases'])):
        c = lambda k: g['c%d_%s' % (i, k)]
        X, Xc, Y, Yc, ls, var, die = (c(k) for k in ('X', 'Xc', 'Y', 'Yc', 'ls', 'var', 'die'))
        N, Din = X.shape[1:]
        Nc, P, ns = Xc.shape[1], Y.shape[2], die.shape[0]
        cls = kernel_class(c('kernel'))
        X_var, Xc_var, Yc_var = Variable(shape=(N, Din)), Variable(shape=(Nc, Din)), Variable(shape=(Nc, P))
        kern = cls(input_dim=Din, ARD=True)
        gp = GaussianProcess.define_variable(X=X_var, kernel=kern, shape=(N, P),
                                             rand_gen=MockMXNetRandomGenerator(T(die.flatten()))).factor
        kp = {kern.name + '_lengthscale': ls, kern.name + '_variance': var}
        variables = {gp.X.uuid: T(X), gp.random_variable.uuid: T(Y)}
        variables.update({getattr(gp, n).uuid: T(v) for n, v in kp.items()})
        yield 'c%d gp log_pdf' % i, gp.log_pdf(F=F, variables=variables).cpu().numpy(), c('gp_log_pdf')
        variables1 = {gp.X.uuid: T(X[:1])}
        variables1.update({getattr(gp, n).uuid: T(v[:1]) for n, v in kp.items()})
        yield 'c%d gp draw' % i, gp.draw_samples(F=F, variables=variables1, num_samples=ns).cpu().numpy(), c('gp_draw')
        kern2 = cls(input_dim=Din, ARD=True)
        cgp = ConditionalGaussianProcess.define_variable(X=X_var, X_cond=Xc_var, Y_cond=Yc_var, kernel=kern2,
                                                         shape=(N, P),
                                                         rand_gen=MockMXNetRandomGenerator(T(die.flatten()))).factor
        variables = {cgp.X.uuid: T(X), cgp.X_cond.uuid: T(Xc), cgp.Y_cond.uuid: T(Yc), cgp.random_variable.uuid: T(Y)}
        variables.update({getattr(cgp, n).uuid: T(v) for n, v in kp.items()})
        yield 'c%d cgp log_pdf' % i, cgp.log_pdf(F=F, variables=variables).cpu().numpy(), c('cgp_log_pdf')
        variables1 = {cgp.X.uuid: T(X[:1]), cgp.X_cond.uuid: T(Xc[:1]), cgp.Y_cond.uuid: T(Yc[:1])}
        variables1.update({getattr(cgp, n).uuid: T(v[:1]) for n, v in kp.items()})
        yield 'c%d cgp draw' % i, cgp.draw_samples(F=F, variables=variables1, num_samples=ns).cpu().numpy(), c('cgp_draw')


def run_svgp_hetero_case(mf, g, i, device):
    """Case i of svgp_hetero.npz: SVGP with one noise variance per data point (svgp_regression.py:61-67)."""
    from mxfusion_b200.components.variables import PositiveTransformation
    from mxfusion_b200.components.distributions.gp.kernels import RBF
    from mxfusion_b200.modules.gp_modules import SVGPRegression
    from mxfusion_b200.inference import GradBasedInference, MAP
    c = lambda k: g['case%d_%s' % (i, k)]
    X, Y, Z, noise_var = c('X'), c('Y'), c('Z'), c('noise_var')
    N, Din = X.shape
    M, P = Z.shape[0], Y.shape[1]
    m = mf.Model()
    m.N = mf.Variable()
    m.X = mf.Variable(shape=(m.N, Din))
    m.Z = mf.Variable(shape=(M, Din), initial_value=Z)
    m.noise_var = mf.Variable(shape=noise_var.shape, transformation=PositiveTransformation(), initial_value=noise_var)
    kernel = RBF(input_dim=Din, ARD=True, variance=c('variance'), lengthscale=c('lengthscale'))
    m.Y = SVGPRegression.define_variable(X=m.X, kernel=kernel, noise_var=m.noise_var, inducing_inputs=m.Z,
                                         shape=(m.N, P))
    gp = m.Y.factor
    gp.svgp_log_pdf.jitter = 1e-8
    infr = GradBasedInference(MAP(model=m, observed=[m.X, m.Y]), context=device)
    infr.initialize(X=X.shape, Y=Y.shape)
    post = gp._extra_graphs[0]
    infr.params[post.qU_mean] = c('qU_mean')
    infr.params[post.qU_cov_W] = c('qU_cov_W')
    infr.params[post.qU_cov_diag] = c('qU_cov_diag')
    infr.params.gflat.zero_()
    loss, loss_g = infr.create_executor()(None, torch.tensor(X, device=device), torch.tensor(Y, device=device))
    loss_g.backward()
    grads = dict(Z=param_grad(infr, m.Z), noise_var=param_grad(infr, m.noise_var),
                 qU_mean=param_grad(infr, post.qU_mean), qU_cov_W=param_grad(infr, post.qU_cov_W),
                 qU_cov_diag=param_grad(infr, post.qU_cov_diag), lengthscale=param_grad(infr, kernel.lengthscale),
                 variance=param_grad(infr, kernel.variance))
    return float(loss), grads


def run_sampling_prediction(mf, g, module, device):
    """Rebuilds the sampling-prediction scenario of sampling_prediction.npz; returns {(noise_free, diag): samples}."""
    from mxfusion_b200.components.variables import PositiveTransformation
    from mxfusion_b200.components.distributions import MockMXNetRandomGenerator
    from mxfusion_b200.components.distributions.gp.kernels import RBF
    from mxfusion_b200.modules.gp_modules import (GPRegression, SVGPRegression, SparseGPRegression,
                                                  GPRegressionSamplingPrediction, SVGPRegressionSamplingPrediction,
                                                  SparseGPRegressionSamplingPrediction)
    from mxfusion_b200.inference import Inference, MAP, TransferInference, ModulePredictionAlgorithm
    X, Y, Z, Xt, die = g['X'], g['Y'], g['Z'], g['Xt'], g['die']
    N, Din = X.shape
    M, P, ns = Z.shape[0], Y.shape[1], die.shape[0]
    m = mf.Model()
    m.N = mf.Variable()
    m.X = mf.Variable(shape=(m.N, Din))
    m.noise_var = mf.Variable(transformation=PositiveTransformation(), initial_value=g['noise_var'])
    kernel = RBF(input_dim=Din, ARD=True, variance=g['variance'], lengthscale=g['lengthscale'])
    if module == 'gp':
        m.Y = GPRegression.define_variable(X=m.X, kernel=kernel, noise_var=m.noise_var, shape=(m.N, P))
    else:
        m.Z = mf.Variable(shape=(M, Din), initial_value=Z)
        cls = SVGPRegression if module == 'svgp' else SparseGPRegression
        m.Y = cls.define_variable(X=m.X, kernel=kernel, noise_var=m.noise_var, inducing_inputs=m.Z, shape=(m.N, P))
        (m.Y.factor.svgp_log_pdf if module == 'svgp' else m.Y.factor.sgp_log_pdf).jitter = 1e-8
    gp = m.Y.factor
    infr = Inference(MAP(model=m, observed=[m.X, m.Y]), context=device)
    infr.initialize(X=X.shape, Y=Y.shape)
    if module == 'svgp':
        post = gp._extra_graphs[0]
        infr.params[post.qU_mean] = g['qU_mean']
        infr.params[post.qU_cov_W] = g['qU_cov_W']
        infr.params[post.qU_cov_diag] = g['qU_cov_diag']
    infr.run(X=X, Y=Y)
    cls = {'gp': GPRegressionSamplingPrediction, 'svgp': SVGPRegressionSamplingPrediction,
           'sparsegp': SparseGPRegressionSamplingPrediction}[module]
    name = {'gp': 'gp_predict', 'svgp': 'svgp_predict', 'sparsegp': 'sgp_predict'}[module]
    out = {}
    for noise_free in (True, False):
        for diag in (True, False):
            alg = cls(gp._module_graph, gp._extra_graphs[0], [gp._module_graph.X],
                      rand_gen=MockMXNetRandomGenerator(torch.tensor(die.flatten(), device=device)))
            alg.noise_free, alg.diagonal_variance, alg.jitter = noise_free, diag, 1e-6
            gp.attach_prediction_algorithms(targets=gp.output_names, conditionals=gp.input_names, algorithm=alg,
                                            alg_name=name)
            infr2 = TransferInference(ModulePredictionAlgorithm(m, observed=[m.X], target_variables=[m.Y],
                                                                num_samples=ns),
                                      infr_params=infr.params, context=device)
            with torch.no_grad():
                out[(noise_free, diag)] = infr2.run(X=Xt)[0].cpu().numpy()
    return out


def run_vpfs_toy(mf, g, device):
    """VariationalPosteriorForwardSampling on the conjugate toy model of vpfs_toy.npz -> (samples of y, samples of mu)."""
    from mxfusion_b200.components.variables import PositiveTransformation
    from mxfusion_b200.components.distributions import Normal, MockMXNetRandomGenerator
    from mxfusion_b200.inference import (GradBasedInference, StochasticVariationalInference, create_Gaussian_meanfield,
                                         VariationalPosteriorForwardSampling)
    y = g['y']
    N, S = y.shape[0], int(g['S'])
    t = lambda a: torch.tensor(np.atleast_1d(a), dtype=torch.float64)
    m = mf.Model()
    m.mu = Normal.define_variable(mean=t(0.), variance=t(4.), shape=(1,))
    m.s2 = mf.Variable(shape=(1,), transformation=PositiveTransformation(), initial_value=t(g['s2']))
    m.y = Normal.define_variable(mean=m.mu, variance=m.s2, shape=(N, 1),
                                 rand_gen=MockMXNetRandomGenerator(torch.tensor(g['eps_y'].flatten(), device=device)))
    q = create_Gaussian_meanfield(model=m, observed=[m.y])
    q.mu.factor._rand_gen = MockMXNetRandomGenerator(torch.tensor(g['eps_mu'].flatten(), device=device))
    alg = StochasticVariationalInference(num_samples=S, model=m, posterior=q, observed=[m.y])
    infr = GradBasedInference(inference_algorithm=alg, context=device)
    infr.initialize(y=y.shape)
    infr.params[q.mu.factor.mean] = t(g['q_mean'])
    infr.params[q.mu.factor.variance] = t(g['q_var'])
    infr2 = VariationalPosteriorForwardSampling(S, [], infr, [m.y, m.mu], context=device)
    with torch.no_grad():
        res = infr2.run()
    return res[0].cpu().numpy(), res[1].cpu().numpy()
