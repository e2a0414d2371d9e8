import sys
sys.path.insert(0,'/root/repo')
exec(open('/root/repo/scratch/ml_exp.py').read().split("def run(")[0])
kk = np.ascontiguousarray(k1[775:780]); P1 = PL(kk)
for eps in (1e-7, 1e-8, 1e-9, 1e-10):
    I = R.Imn(PL, kk, 0, 1, epsrel=eps); J = R.Jmn(PL, kk, 0, 1, epsrel=min(eps,1e-9))
    print(eps, (2*I + 6*kk**2*J*P1 - com[1][775:780])/np.maximum(np.abs(com[1][775:780]),P1))
Ie,_ = E.ik(1, kk, tabL, tabL, 0, qmax=1e5); Je = E.linear(0, 1, kk, 1e-5, 1e5, tabL)
print('emul', (2*Ie + 6*kk**2*Je*P1 - com[1][775:780])/np.maximum(np.abs(com[1][775:780]),P1))
print('I rel emul vs 1e-10', Ie/I-1, 'J', Je/J-1)
for eps in (1e-4, 1e-5,1e-6,1e-7,1e-8,1e-9,1e-11):
    print('J01 eps', eps, R.Jmn(PL, kk, 0, 1, epsrel=eps)/J-1)
