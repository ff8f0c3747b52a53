# /* Copyright (C) 2001
#  * Housemarque Oy
#  * http://www.housemarque.com
#  *
#  * Distributed under the Boost Software License, Version 1.0. (See
#  * accompanying file LICENSE_1_0.txt or copy at
#  * http://www.boost.org/LICENSE_1_0.txt)
#  */
#
# /* Revised by Paul Mensonides (2002) */
# /* Revised by Edward Diener (2011,2013) */
#
# /* See http://www.boost.org for most recent version. */
#
# ifndef BOOST_PREPROCESSOR_TUPLE_HPP
# define BOOST_PREPROCESSOR_TUPLE_HPP
#
# include <libint2/boost/preprocessor/tuple/eat.hpp>
# include <libint2/boost/preprocessor/tuple/elem.hpp>
# include <libint2/boost/preprocessor/tuple/enum.hpp>
# include <libint2/boost/preprocessor/tuple/insert.hpp>
# include <libint2/boost/preprocessor/tuple/pop_back.hpp>
# include <libint2/boost/preprocessor/tuple/pop_front.hpp>
# include <libint2/boost/preprocessor/tuple/push_back.hpp>
# include <libint2/boost/preprocessor/tuple/push_front.hpp>
# include <libint2/boost/preprocessor/tuple/rem.hpp>
# include <libint2/boost/preprocessor/tuple/remove.hpp>
# include <libint2/boost/preprocessor/tuple/replace.hpp>
# include <libint2/boost/preprocessor/tuple/reverse.hpp>
# include <libint2/boost/preprocessor/tuple/size.hpp>
# include <libint2/boost/preprocessor/tuple/to_array.hpp>
# include <libint2/boost/preprocessor/tuple/to_list.hpp>
# include <libint2/boost/preprocessor/tuple/to_seq.hpp>
#
# endif
