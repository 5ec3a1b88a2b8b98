import sys,json
for l in open(sys.argv[1]):
    if not l.startswith('{'): continue
    d=json.loads(l)
    print(d['config'][14:22], d['systems'], 'v',d['variant'], 'qr %.3f ldiv %.3f fused %.3f  qrf %.2f ldf %.2f tot %.2f'%(d['qr_ms'],d['ldiv_ms'],d['fused_solve_ms'],d['qr_frac'],d['ldiv_frac'],d['qr_plus_ldiv_frac']))
